/* intra_kernels.cuh — intra pictures (and any picture with more intra than inter macroblocks): the reconstruction
 * wavefront as BANDS of rows.
 *
 * Intra prediction of macroblock (x, y) reads the unfiltered samples of its neighbours A (x-1, y), D, B, C (x-1..x+1, y-1)
 * (reference: decode order edge264_slice.c:1651-1849 with the predictors of edge264_intra.c:291-765): rows can run in
 * parallel two macroblocks apart.  Round 1 ran one warp per row with a global flag, a gpu-scope fence (which also
 * invalidates L1) and reloads of the record, the residual and the row above from L2 inside every step: 8.7 us per
 * macroblock step, 2.2 ms per 1080p I picture.  Here a block owns IR_ROWS consecutive rows (two warps each: one
 * inverse-transforms up to two macroblocks ahead, one predicts); the bottom
 * sample row of every finished macroblock goes to the row below through a shared-memory ring with done/taken counters
 * (block-scope fences); records arrive two macroblocks ahead, coefficient runs one ahead (cp.async.bulk on an mbarrier) and
 * are inverse-transformed BEFORE the row waits for its neighbours, so a step is the prediction itself plus a
 * shared-memory hand-shake and the residual never travels through global memory.  Only the band's last row talks to the next block through global memory, with
 * a progress counter published every IR_CHUNK macroblocks.  Inter macroblocks inside such a picture were reconstructed
 * by the inter kernel before: their row warp only forwards their border samples. */
#pragma once
#include "recon_kernels.cuh"
#include "deblock_kernels.cuh"

#define IR_ROWS 8
#define IR_RING 16
#define IR_CHUNK 8
#define IR_SLOTS 4      /* macroblocks the transform warp may be ahead of the prediction warp (residual / tile slots, coefficient buffers) */
#define IR_RECS 8       /* record ring of a row */
struct __align__(16) IntraRowsSmem {
	WarpSmem ws[IR_ROWS][IR_SLOTS];                    /* macroblocks x and x+1 of every row: residual from the row's transform warp, tile of its prediction warp */
	uint4 recs[IR_ROWS][IR_RECS][12];           /* records of macroblocks x-4 .. x+3 of every row */
	uint32_t ring[IR_ROWS][IR_RING][8];         /* bottom sample row of a finished macroblock: luma (4 words), Cb (2), Cr (2) */
	int16_t coef[IR_ROWS][IR_SLOTS][RES_COEF_MAX];   /* coefficient runs of macroblocks x .. x+3 of every row (cp.async.bulk) */
	unsigned long long bars[IR_ROWS][IR_SLOTS];
	uint4 drecs[IR_ROWS][3][12];                /* deblocking digests of intra-only pictures: current, left and top record */
	E264DbkMb ddg[IR_ROWS];
	int done[IR_ROWS];                          /* macroblocks a row has put into its ring */
	int taken[IR_ROWS];                         /* the macroblock a row is working on: ring entries before it (minus one) are free */
	int ready[IR_ROWS];                         /* macroblocks whose residual (and record) the transform warp has delivered */
	int used[IR_ROWS];                          /* macroblocks the prediction warp is through with */
	int band;
};

/* The row's TRANSFORM warp: records four ahead (through a register, so the warp never waits for the load it has just
 * issued), coefficient runs three ahead (cp.async.bulk; DRAM latency is longer than one macroblock of this warp's work),
 * dequantisation and inverse transforms into the residual of slot x % IR_SLOTS, at most IR_SLOTS - 1 macroblocks ahead of
 * the prediction warp.  In a picture without inter macroblocks it also derives the row's deblocking digests (no inter
 * kernel runs that would). */
__device__ __forceinline__ void intra_row_transform(const PicJob &J, IntraRowsSmem &sm, int band, int w, int lane, unsigned &parbits) {
	const int W = J.w_mbs, H = J.h_mbs, nmb = W * H;
	const int mby = band * IR_ROWS + w;
	if (mby >= H) return;
	const E264MbRec *rowrecs = J.recs + (size_t)mby * W;
	const volatile unsigned *errp = J.err;
	volatile int *used = sm.used + w, *ready = sm.ready + w;
	const bool digests = J.dbk != nullptr && J.n_intra == nmb;
	constexpr int AHEAD = IR_SLOTS - 1;          /* coefficient runs in flight ahead of the macroblock being transformed */
	/* prologue: records 0 .. AHEAD in shared memory, record AHEAD + 1 in the register */
	for (int x = 0; x <= AHEAD && x < W; x++) if (lane < 12) sm.recs[w][x % IR_RECS][lane] = __ldg((const uint4 *)(rowrecs + x) + lane);
	uint4 recreg = make_uint4(0, 0, 0, 0);
	if (AHEAD + 1 < W && lane < 12) recreg = __ldg((const uint4 *)(rowrecs + AHEAD + 1) + lane);
	__syncwarp();
	auto coef_issue = [&](int x) -> bool {       /* true: a copy is in flight into buffer x % IR_SLOTS */
		const E264MbRec *rr = (const E264MbRec *)sm.recs[w][x % IR_RECS];
		if (rr->kind == MBK_INTER || rr->kind == MBK_IPCM || rr->coded == 0) return false;
		if (lane == 0) tma_bulk_g2s(sm.coef[w][x % IR_SLOTS], J.coefs + rr->coef_off, (unsigned)rec_coef_count(rr) * 2u, &sm.bars[w][x % IR_SLOTS]);
		return true;
	};
	unsigned pending = 0;                        /* bit x % IR_SLOTS: macroblock x has a copy in flight */
	RT_DECL
	for (int x = 0; x < AHEAD && x < W; x++) if (coef_issue(x)) pending |= 1u << (x % IR_SLOTS);
#pragma unroll 1
	for (int mbx = 0; mbx < W; mbx++) {
		const int sl = mbx % IR_SLOTS;
		/* residual slot sl is free once the prediction warp is through with macroblock mbx - IR_SLOTS; the record slot written
		 * below belonged to macroblock mbx + AHEAD + 1 - IR_RECS, older still */
		if (lane == 0) {
			unsigned spins = 0; bool bad = false;
			while (*used < mbx - IR_SLOTS + 1 && !bad) { if ((++spins & 1023) == 0) bad = *errp != 0 || spins > (1u << 26); }
			if (bad) atomicExch(J.err, 1u);
			__threadfence_block();
		}
		__syncwarp();
		RT_MARK(0)      /* waiting for a free slot */
		const E264MbRec *r = (const E264MbRec *)sm.recs[w][mbx % IR_RECS];
		WarpSmem *ws = &sm.ws[w][sl];
		/* the record fetched one iteration ago goes to shared memory, the next one is requested */
		if (mbx + AHEAD + 1 < W && lane < 12) sm.recs[w][(mbx + AHEAD + 1) % IR_RECS][lane] = recreg;
		if (mbx + AHEAD + 2 < W && lane < 12) recreg = __ldg((const uint4 *)(rowrecs + mbx + AHEAD + 2) + lane);
		__syncwarp();
		if (mbx + AHEAD < W) { const int nb = (mbx + AHEAD) % IR_SLOTS; pending &= ~(1u << nb); if (coef_issue(mbx + AHEAD)) pending |= 1u << nb; }
		const int kind = r->kind;
		RT_MARK(1)      /* records, coefficient copy issued */
		if ((pending >> sl) & 1) {
			if (!mbar_wait(&sm.bars[w][sl], (parbits >> sl) & 1)) { if (lane == 0) atomicExch(J.err, 3u); }
			parbits ^= 1u << sl;
			RT_MARK(2)  /* waiting for the coefficient run */
			residual_stage(ws, r, J.slices + r->slice_idx, sm.coef[w][sl], lane);
		} else if (kind != MBK_INTER && kind != MBK_IPCM) {
			((uint4 *)ws->res)[lane] = make_uint4(0, 0, 0, 0);
			if (lane < 16) ((uint4 *)ws->res)[32 + lane] = make_uint4(0, 0, 0, 0);
		}
		__syncwarp();
		if (lane == 0) { __threadfence_block(); *ready = mbx + 1; }
		RT_MARK(3)      /* inverse transforms */
		if (digests) dbk_digest_mb(J, sm.drecs[w], &sm.ddg[w], mby * W + mbx, lane);
		RT_MARK(4)      /* deblocking digest */
	}
	RT_FLUSH(0)
}

/* The row's PREDICTION warp: waits for the residual, for the row above (top-right neighbour finished) and for room in
 * its ring, predicts into the tile of slot x & 1, stores, hands the bottom sample row down. */
__device__ __forceinline__ void intra_row_walk(const PicJob &J, IntraRowsSmem &sm, int band, int w, int lane) {
	const int W = J.w_mbs, H = J.h_mbs, nmb = W * H;
	const int lrow = w, mby = band * IR_ROWS + w;
	if (mby >= H) return;
	const bool from_global = lrow == 0 && mby > 0;
	const bool from_ring = lrow > 0;
	const bool to_ring = lrow + 1 < IR_ROWS && mby + 1 < H;
	const bool to_global = lrow + 1 == IR_ROWS && mby + 1 < H;
	volatile unsigned *prog = J.flags + nmb + 2 * H;          /* intra row progress (third counter array) */
	const volatile unsigned *errp = J.err;
	const unsigned base = J.epoch * 2048u;
	volatile int *done_in = sm.done + (lrow > 0 ? lrow - 1 : 0), *done_out = sm.done + lrow;
	volatile int *taken_me = sm.taken + lrow, *taken_next = sm.taken + (lrow + 1 < IR_ROWS ? lrow + 1 : lrow);
	volatile int *ready = sm.ready + w, *used = sm.used + w;
	uint32_t (*ring_in)[8] = sm.ring[lrow > 0 ? lrow - 1 : 0];
	uint32_t (*ring_out)[8] = sm.ring[lrow];
	uint8_t *dst = J.frames + (size_t)J.dst_slot * J.frame_bytes;
	const int cpl = J.stride_c >> 1;
	int avail = 0;                                            /* macroblocks of the row above the band known to be stored */
	RT_DECL
#pragma unroll 1
	for (int mbx = 0; mbx < W; mbx++) {
		uint8_t *Y = dst + (size_t)(mby * 16) * J.stride_y + mbx * 16;
		uint8_t *C = dst + J.plane_y + (size_t)(mby * 8) * J.stride_c + mbx * 8;
		/* ---- wait: residual delivered; row above through with macroblocks up to x+1 (top-right neighbour); room in our ring ---- */
		if (lane == 0) {
			unsigned spins = 0; bool bad = false;
			while (*ready <= mbx && !bad) { if ((++spins & 1023) == 0) bad = *errp != 0 || spins > (1u << 26); }
			RT_MARK(0)  /* waiting for the residual */
			const int need = min(mbx + 2, W);
			if (from_ring) {
				while (*done_in < need && !bad) { if ((++spins & 1023) == 0) bad = *errp != 0 || spins > (1u << 26); }
			} else if (from_global && avail < need) {
				const unsigned target = base + (unsigned)need;
				unsigned v = prog[mby - 1];
				while ((int)(v - target) < 0 && !bad) { __nanosleep(40); if ((++spins & 255) == 0) bad = *errp != 0 || spins > (1u << 22); v = prog[mby - 1]; }
				__threadfence();
				avail = bad ? W : (int)(v - base);
			}
			RT_MARK(1)  /* waiting for the row above */
			if (to_ring) while (*taken_next <= mbx - IR_RING + 1 && !bad)     /* entry mbx - RING is still the corner sample of the row below's macroblock mbx - RING + 1 */ { if ((++spins & 1023) == 0) bad = *errp != 0 || spins > (1u << 26); }
			if (bad) atomicExch(J.err, 1u);
			__threadfence_block();
			*taken_me = mbx;
		}
		__syncwarp();
		const E264MbRec *r = (const E264MbRec *)sm.recs[w][mbx % IR_RECS];
		const int kind = r->kind;
		WarpSmem *ws = &sm.ws[w][mbx % IR_SLOTS];
		/* ---- the row above into the tile ---- */
		if (from_ring) {
			if (lane < 4) *(uint32_t *)&YT(4 * lane, -1) = ring_in[mbx % IR_RING][lane];
			else if (lane < 6) { if (mbx + 1 < W) *(uint32_t *)&YT(16 + 4 * (lane - 4), -1) = ring_in[(mbx + 1) % IR_RING][lane - 4]; }
			else if (lane == 6) { if (mbx > 0) YT(-1, -1) = (uint8_t)(ring_in[(mbx - 1) % IR_RING][3] >> 24); }
			else if (lane >= 8 && lane < 12) *(uint32_t *)&CT((lane - 8) >> 1, 4 * (lane & 1), -1) = ring_in[mbx % IR_RING][lane - 4];
			else if (lane == 12) { if (mbx > 0) CT(0, -1, -1) = (uint8_t)(ring_in[(mbx - 1) % IR_RING][5] >> 24); }
			else if (lane == 13) { if (mbx > 0) CT(1, -1, -1) = (uint8_t)(ring_in[(mbx - 1) % IR_RING][7] >> 24); }
		} else if (from_global) {
			const int x = lane - 1;   /* -1..23 */
			if (x < 24 && (x >= 0 || mbx > 0) && (x < 16 || mbx < W - 1)) YT(x, -1) = __ldcg(Y - J.stride_y + x);
			if (lane < 18) { const int pl = lane / 9, cx = lane % 9 - 1; if (cx >= 0 || mbx > 0) CT(pl, cx, -1) = __ldcg(C + pl * cpl - J.stride_c + cx); }
		}
		__syncwarp();
		RT_MARK(2)      /* ring room, hand-shake, row above into the tile */
		/* ---- the macroblock ---- */
		if (kind == MBK_INTER) {   /* reconstructed by the inter kernel: fetch its right column and bottom row for the neighbours to come */
			if (lane < 16) YT(15, lane) = __ldcg(Y + (size_t)lane * J.stride_y + 15);
			else { const int j = lane - 16; CT(j >> 3, 7, j & 7) = __ldcg(C + (j >> 3) * cpl + (size_t)(j & 7) * J.stride_c + 7); }
			if (to_ring) {
				if (lane < 4) ring_out[mbx % IR_RING][lane] = __ldcg((const uint32_t *)(Y + (size_t)15 * J.stride_y) + lane);
				else if (lane < 8) ring_out[mbx % IR_RING][lane] = __ldcg((const uint32_t *)(C + ((lane - 4) >> 1) * cpl + (size_t)7 * J.stride_c) + (lane & 1));
			}
		} else {
			if (kind == MBK_IPCM) {
				const uint8_t *s = (const uint8_t *)(J.coefs + r->coef_off);
				if (lane < 16) *(uint4 *)&YT(0, lane) = __ldg((const uint4 *)s + lane);
				else { const int j = lane - 16; *(uint2 *)&CT(j >> 3, 0, j & 7) = __ldg((const uint2 *)(s + 256) + j); }
				__syncwarp();
			} else {
				intra_luma(ws, r, lane);
				RT_MARK(3)  /* luma prediction */
				intra_chroma(ws, r, lane);
				RT_MARK(4)  /* chroma prediction */
			}
			store_mb(ws, J, Y, C, lane);
			if (to_ring) {
				if (lane < 4) ring_out[mbx % IR_RING][lane] = *(const uint32_t *)&YT(4 * lane, 15);
				else if (lane < 8) ring_out[mbx % IR_RING][lane] = *(const uint32_t *)&CT((lane - 4) >> 1, 4 * (lane & 1), 7);
			}
		}
		__syncwarp();
		RT_MARK(5)      /* stores */
		/* ---- publish: shared-memory counter for the row below in this band; global counter for the next band, one chunk late ---- */
		if (lane == 0) {
			if (to_ring) { __threadfence_block(); *done_out = mbx + 1; }
			if (to_global && mbx > 0 && mbx % IR_CHUNK == 0) { __threadfence(); prog[mby] = base + (unsigned)mbx; }   /* covers macroblocks 0..mbx-1, stored in earlier iterations */
		}
		/* right-most column becomes the next macroblock's left neighbour (the other slot's tile) */
		{
			WarpSmem *wn = &sm.ws[w][(mbx + 1) % IR_SLOTS];
			if (lane < 16) wn->ytile[(lane + 1) * YT_STRIDE + 15] = YT(15, lane);
			else if (lane < 24) wn->ctile[0][(lane - 16 + 1) * CT_STRIDE + 7] = CT(0, 7, lane - 16);
			if (lane < 8) wn->ctile[1][(lane + 1) * CT_STRIDE + 7] = CT(1, 7, lane);
		}
		__syncwarp();
		if (lane == 0) { __threadfence_block(); *used = mbx + 1; }
		avail = __shfl_sync(0xffffffffu, avail, 0);
		RT_MARK(6)      /* publication, column hand-over */
	}
	RT_FLUSH(8)
	__syncwarp();
	if (lane == 0) {
		*taken_me = W + IR_RING;       /* nothing of the ring above is needed any more */
		if (to_global) { __threadfence(); prog[mby] = base + (unsigned)W; }
	}
}

__global__ void __launch_bounds__(IR_ROWS * 64) e264_intra_rows_kernel(PicJob J) {
	TraceScope trace_(J, 2);
	reset_next_tickets(J);
	extern __shared__ __align__(16) unsigned char ir_smem_raw[];
	IntraRowsSmem &sm = *(IntraRowsSmem *)ir_smem_raw;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, w = wid & (IR_ROWS - 1);
	const bool transform = wid >= IR_ROWS;
	const int bands = (J.h_mbs + IR_ROWS - 1) / IR_ROWS;
	unsigned parbits = 0;
	if (transform && lane == 0) {
		for (int k = 0; k < IR_SLOTS; k++) mbar_init(&sm.bars[w][k], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	/* bands in dispatch order: a band only waits for bands drawn before it */
	for (;;) {
		__syncthreads();
		if (threadIdx.x < IR_ROWS) { sm.done[threadIdx.x] = 0; sm.taken[threadIdx.x] = 0; sm.ready[threadIdx.x] = 0; sm.used[threadIdx.x] = 0; }
		if (threadIdx.x == 0) sm.band = (int)atomicAdd(J.tickets + 2, 1u);
		__syncthreads();
		const int band = sm.band;
		if (band >= bands) break;
		if (transform) intra_row_transform(J, sm, band, w, lane, parbits);
		else intra_row_walk(J, sm, band, w, lane);
	}
}


/* ---- intra macroblocks of P and B pictures ----
 * The parser lists them in raster order (E264Staging.intra_list); warps draw list entries by ticket, so a macroblock
 * only ever waits for entries drawn before its own.  Neighbours: A, D, B, C "reconstructed" flags — inter neighbours were
 * flagged by e264_inter4_kernel (kernel boundary), intra ones are flagged here after a fence.  The coefficient run is
 * fetched (cp.async.bulk) and inverse-transformed while the neighbours finish. */
/* Block size: the kernel's span is the longest chain of dependent intra macroblocks, not its amount of work, and the inter
 * kernel's blocks need a WHOLE SM: a few big blocks (INTRA_WARPS warps, several list entries per warp) leave most SMs to
 * them, where many 4-warp blocks put one small, latency-bound block on nearly every SM. */
#ifndef INTRA_WARPS
#define INTRA_WARPS 16
#endif
struct __align__(16) IntraTkSmem {
	WarpSmem ws[INTRA_WARPS];
	int16_t coef[INTRA_WARPS][RES_COEF_MAX];
	unsigned long long bars[INTRA_WARPS];
};
__global__ void __launch_bounds__(INTRA_WARPS * 32) e264_intra_kernel(PicJob J) {
	TraceScope trace_(J, 2);
	reset_next_tickets(J);
	extern __shared__ __align__(16) unsigned char intra_tk_smem_raw[];
	IntraTkSmem &sm = *(IntraTkSmem *)intra_tk_smem_raw;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	WarpSmem *ws = &sm.ws[w];
	const int nmb = J.w_mbs * J.h_mbs, W = J.w_mbs, cpl = J.stride_c >> 1;
	uint8_t *dst = J.frames + (size_t)J.dst_slot * J.frame_bytes;
	unsigned parity = 0;
	if (lane == 0) {
		mbar_init(&sm.bars[w], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncwarp();
	for (;;) {
		unsigned t = 0;
		if (lane == 0) t = atomicAdd(J.tickets + 2, 1u);
		t = __shfl_sync(0xffffffffu, t, 0);
		if (t >= (unsigned)J.n_intra) break;
		const int mb = (int)__ldg(J.intra_list + t);
		if (mb >= nmb) continue;
		const int mbx = mb % W, mby = mb / W;
		if (lane < 12) ws->rec4[lane] = __ldg((const uint4 *)(J.recs + mb) + lane);
		__syncwarp();
		const E264MbRec *r = (const E264MbRec *)ws->rec4;
		const int kind = r->kind;
		if (kind == MBK_INTER) continue;      /* (the list holds intra macroblocks only) */
		uint8_t *Y = dst + (size_t)(mby * 16) * J.stride_y + mbx * 16;
		uint8_t *C = dst + J.plane_y + (size_t)(mby * 8) * J.stride_c + mbx * 8;
		if (kind == MBK_IPCM) {
			const uint8_t *s = (const uint8_t *)(J.coefs + r->coef_off);
			if (lane < 16) *(uint4 *)&YT(0, lane) = __ldg((const uint4 *)s + lane);
			else { const int j = lane - 16; *(uint2 *)&CT(j >> 3, 0, j & 7) = __ldg((const uint2 *)(s + 256) + j); }
			__syncwarp();
			store_mb(ws, J, Y, C, lane);
		} else {
			const bool coded = r->coded != 0;
			if (coded && lane == 0) tma_bulk_g2s(sm.coef[w], J.coefs + r->coef_off, (unsigned)rec_coef_count(r) * 2u, &sm.bars[w]);
			if (coded) {
				if (!mbar_wait(&sm.bars[w], parity)) { if (lane == 0) atomicExch(J.err, 3u); }
				parity ^= 1;
				residual_stage(ws, r, J.slices + r->slice_idx, sm.coef[w], lane);
			} else {
				((uint4 *)ws->res)[lane] = make_uint4(0, 0, 0, 0);
				if (lane < 16) ((uint4 *)ws->res)[32 + lane] = make_uint4(0, 0, 0, 0);
			}
			{	/* A, D, B, C: one lane each, the four round trips overlap */
				int idx = -1;
				if (lane == 0 && mbx > 0) idx = mb - 1;
				else if (lane == 1 && mby > 0 && mbx > 0) idx = mb - W - 1;
				else if (lane == 2 && mby > 0) idx = mb - W;
				else if (lane == 3 && mby > 0) idx = mbx < W - 1 ? mb - W + 1 : mb - W;
				if (idx >= 0) wait_flag_acquire(J.flags, idx, J.epoch, J.err);
			}
			__syncwarp();
			if (mby > 0) {
				const int x = lane - 1;   /* -1..23 */
				if (x < 24 && (x >= 0 || mbx > 0) && (x < 16 || mbx < W - 1)) YT(x, -1) = __ldcg(Y - J.stride_y + x);
				if (lane < 18) { const int pl = lane / 9, cx = lane % 9 - 1; if (cx >= 0 || mbx > 0) CT(pl, cx, -1) = __ldcg(C + pl * cpl - J.stride_c + cx); }
			}
			if (mbx > 0) {
				if (lane < 16) YT(-1, lane) = __ldcg(Y + (size_t)lane * J.stride_y - 1);
				else { const int j = lane - 16, pl = j >> 3, row = j & 7; CT(pl, -1, row) = __ldcg(C + pl * cpl + (size_t)row * J.stride_c - 1); }
			}
			__syncwarp();
			intra_luma(ws, r, lane);
			intra_chroma(ws, r, lane);
			store_mb(ws, J, Y, C, lane);
		}
		__syncwarp();
		if (lane == 0) { __threadfence(); *(volatile unsigned *)(J.flags + mb) = J.epoch; }
		__syncwarp();
	}
}
