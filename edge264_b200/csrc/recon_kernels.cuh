/* recon_kernels.cuh — sm_100a device code of the pixel-reconstruction path.
 *
 * One warp per macroblock (per macroblock row where a wavefront runs).  Up to four kernels per picture:
 *   e264_residual_kernel  inverse quantisation + 4x4/8x8 inverse transforms + DC transforms
 *                         (reference edge264_residual.c:108-538); coefficient runs staged by cp.async.bulk;
 *   e264_inter_kernel     6-tap / bilinear motion compensation with default / explicit / implicit weighting
 *                         (edge264_inter.c:416-1251); reference windows staged by cp.async.bulk.tensor;
 *   e264_intra_kernel     intra prediction of every mode and I_PCM (edge264_intra.c:291-765, slice.c:886-939);
 *   e264_deblock_kernel   boundary strengths and the in-loop filter (edge264_deblock.c:284-1123).
 * Inter macroblocks have no dependency inside a picture (tickets); an intra macroblock waits on the "done" flags
 * of its A/B/C/D neighbours; a deblocking row-warp waits on the progress counter of the row above — the same
 * dependencies the reference resolves by decoding in raster order with deblocking one row behind
 * (edge264_slice.c:1809-1826).  Arithmetic is restated from ITU-T H.264 with the reference's observable integer
 * widths; results are bit-exact with the reference decoder (tests/test_gpu_parity.py).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "records.h"
#include "h264_tables.h"

struct PicJob {
	const E264MbRec *recs; const int16_t *coefs; const E264SliceRec *slices;
	uint8_t *frames;
	int frame_bytes, w_mbs, h_mbs, stride_y, stride_c, plane_y, dst_slot, n_slots;
	unsigned *flags;      /* [nmb] "reconstructed" == epoch, then [3][h_mbs] row progress (epoch * 2048 + macroblocks done): deblocking luma, deblocking chroma, intra pictures */
	unsigned epoch;
	unsigned *tickets;    /* [8] zeroed by e264_prepass_kernel: 0 inter, 1 deblock, 2 intra */
	unsigned *err;
	int rows_mode;
	int resid_inter;      /* e264_residual_kernel also transforms inter macroblocks (only the round-1 inter kernel reads them from J.resid) */
	struct E264DbkMb *dbk;        /* [nmb] deblocking digests written by e264_prepass_kernel; NULL = picture is not deblocked */
	const uint32_t *intra_list;   /* addresses of the intra macroblocks in raster order */
	int n_intra;
	const void *tmaps;    /* CUtensorMap[6] over the whole frame pool (x, y, slot): luma boxes 48x{21,13,9}, chroma boxes 32x{9,5,3}; NULL = no TMA */
	int16_t *resid;       /* [nmb][384] residual written by e264_residual_kernel (coded macroblocks only) */
	unsigned long long *trace;   /* measurement only (e264b_replay): [trace_base + kind] = {first block start, last block end} in globaltimer ns; kinds: 0 residual, 1 inter, 2 intra, 3 deblock, 4 prepass */
	int trace_base;
	int phase_slot;              /* index into trace[] of the 10 phase counters (E264B_PHASES builds) */
};

#define WARPS_PER_BLOCK 4
#define YT_STRIDE 48     /* luma tile row: [15]=left neighbour, [16..31]=samples, [32..39]=top-right */
#define CT_STRIDE 16     /* chroma tile row: [7]=left neighbour, [8..15]=samples */
/* motion-compensation window buffer (one TMA destination set): luma box 48 x 21 at +0, Cb box 32 x 9 at +1024,
 * Cr box 32 x 9 at +1408 — every box starts on a 128-byte boundary as cp.async.bulk.tensor requires.  A box must
 * start on a 16-byte boundary of the picture row (measured: other x coordinates raise an illegal-instruction
 * fault), so the wanted window begins 0..15 bytes into its shared-memory rows. */
#define WIN_STRIDE 48
#define WIN_C_STRIDE 32
#define WIN_CB_OFF 1024
#define WIN_CR_OFF 1408
#define WIN_BYTES 1792

struct __align__(16) WarpSmem {
	uint4 rec4[12];                 /* the macroblock record */
	int16_t res[384];               /* residual: luma y*16+x, then Cb, Cr 8x8 */
	uint8_t ytile[17 * YT_STRIDE];  /* row 0 = samples above the macroblock */
	uint8_t ctile[2][9 * CT_STRIDE];
	int dc[24];                     /* scaled DC: 16 luma (raster over blocks), 4 Cb, 4 Cr */
	uint8_t pt1[256 + 128];         /* second prediction of bi-predicted 8x8 quadrants: luma y*16+x, then Cb, Cr 8x8 */
	int wq[4][3][4];                /* per 8x8 quadrant and component: {mode, w0, w1, offset | log2wd << 16}, see mc_blend */
	union {
		int16_t t8[4 * 64];         /* 8x8 transform transpose buffer */
		int edge[2][28];            /* intra 8x8 filtered reference samples */
	} u;
};

__device__ __forceinline__ int clip255(int v) { return min(max(v, 0), 255); }
__device__ __forceinline__ int sat16(int v) { return min(max(v, -32768), 32767); }
__device__ __forceinline__ int blk_x(int b) { return (b & 1) | ((b >> 1) & 2); }
__device__ __forceinline__ int blk_y(int b) { return ((b >> 1) & 1) | ((b >> 2) & 2); }
__device__ __forceinline__ int blk_z(int x, int y) { return (x & 1) | ((y & 1) << 1) | ((x & 2) << 1) | ((y & 2) << 2); }
__device__ __forceinline__ int norm4(int m, int i, int j) { return h264_norm4x4[m][((i & 1) && (j & 1)) ? 1 : (!(i & 1) && !(j & 1)) ? 0 : 2]; }
__device__ __forceinline__ int norm8(int m, int i, int j) {
	int k;
	if (!(i & 3) && !(j & 3)) k = 0;
	else if ((i & 1) && (j & 1)) k = 1;
	else if ((i & 3) == 2 && (j & 3) == 2) k = 2;
	else if ((!(i & 3) && (j & 1)) || ((i & 1) && !(j & 3))) k = 3;
	else if ((!(i & 3) && (j & 3) == 2) || ((i & 3) == 2 && !(j & 3))) k = 4;
	else k = 5;
	return h264_norm8x8[m][k];
}

/* measurement only: first-start / last-end timestamps of a launch, see e264b_replay */
struct TraceScope {
	unsigned long long *t;
	__device__ __forceinline__ static unsigned long long now() { unsigned long long v; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v)); return v; }
	/* one thread per block: two atomics per block keep the timed replay undisturbed (a block's first warp starts it, its exit is within one macroblock of the block's end) */
	__device__ __forceinline__ TraceScope(const PicJob &J, int kind) { t = J.trace ? J.trace + 2 * (J.trace_base + kind) : nullptr; if (t && threadIdx.x == 0) atomicMin(t, now()); }
	__device__ __forceinline__ ~TraceScope() { if (t && threadIdx.x == 0) atomicMax(t + 1, now()); }
};

/* -DE264B_PHASES (measurement builds only): per-phase clock accumulation in the inter kernel; lane 0 of each warp
 * adds its totals to J.trace[PHASE_SLOT + i] at exit.  Phases: 0 ticket+record, 1 residual fetch, 2 rect list,
 * 3 window issue, 4 TMA wait, 5 luma filter, 6 chroma filter, 7 blend+residual add, 8 store. */
#ifdef E264B_PHASES
#define PH_DECL long long ph_acc[10] = {0,0,0,0,0,0,0,0,0,0}; long long ph_t = clock64();
#define PH(i) do { long long c_ = clock64(); ph_acc[i] += c_ - ph_t; ph_t = c_; } while (0)
#define PH_ARGS , long long *ph_acc, long long &ph_t
#define PH_PASS , ph_acc, ph_t
#define PH_FLUSH(J) do { if ((J).trace && (threadIdx.x & 31) == 0) for (int i_ = 0; i_ < 10; i_++) atomicAdd((J).trace + (J).phase_slot + i_, (unsigned long long)ph_acc[i_]); } while (0)
#else
#define PH_DECL
#define PH(i)
#define PH_ARGS
#define PH_PASS
#define PH_FLUSH(J)
#endif

/* spin until flags[idx] == epoch (lane 0).  Bounded so that a bug cannot hang the GPU: gives up after ~0.2 s of SM
 * clocks, or at once when another warp has already raised the error flag. */
__device__ __forceinline__ bool wait_flag(const unsigned *flags, int idx, unsigned epoch, unsigned *err) {
	const volatile unsigned *f = flags + idx;
	unsigned spins = 0;
	long long t0 = 0;
	while (*f != epoch) {
		__nanosleep(64);
		if ((++spins & 63) == 0) {
			if (*(const volatile unsigned *)err) return false;
			if (t0 == 0) t0 = clock64();
			else if (clock64() - t0 > 400000000ll) { atomicExch(err, 1u); return false; }
		}
	}
	return true;
}

/* spin until *p - need >= 0 (counters carry the picture epoch in their upper bits), bounded like wait_flag */
__device__ __forceinline__ bool wait_progress(const unsigned *p, unsigned need, unsigned *err) {
	const volatile unsigned *f = p;
	unsigned spins = 0;
	long long t0 = 0;
	while ((int)(*f - need) < 0) {
		__nanosleep(64);
		if ((++spins & 63) == 0) {
			if (*(const volatile unsigned *)err) return false;
			if (t0 == 0) t0 = clock64();
			else if (clock64() - t0 > 400000000ll) { atomicExch(err, 1u); return false; }
		}
	}
	return true;
}

/* ---- shared-memory barrier + TMA helpers (sm_90+ PTX; SASS: SYNCS.*, UBLKCP, UTMALDG) ---- */
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, void *bar) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
/* wait for the phase with the given parity; gives up after ~10 ms of SM clocks (a window arrives in
 * microseconds), false = gave up (the caller raises the error flag).  One loop, one exit: ptxas reconverges the
 * warp behind it, which the __syncwarp()s that follow rely on (they compile to NOPs where the compiler has proved
 * convergence). */
__device__ __forceinline__ bool mbar_wait(void *bar, unsigned parity) {
	unsigned done = 0;
	const long long t0 = clock64();
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	} while (!done && clock64() - t0 < 20000000ll);
	return done != 0;
}

/* ------------------------------------------------------------------------------------------ */
/* residual                                                                                     */
/* ------------------------------------------------------------------------------------------ */
/* 4x4 inverse transform of 16 levels at c (16-byte aligned) -> residual written to dst[y*dstride+x] */
__device__ __noinline__ void idct4x4(const int16_t *c, bool have_levels, const uint8_t *scaling, int qp, bool dc_override, int dc, int16_t *dst, int dstride) {
	int d[16];
	if (have_levels) {
		uint4 a = *(const uint4 *)c, b = *((const uint4 *)c + 1);   /* c points into the TMA-staged shared copy */
		int16_t lv[16];
		*(uint4 *)lv = a; *(uint4 *)(lv + 8) = b;
		int sh = qp / 6, m = qp - sh * 6;
#pragma unroll
		for (int i = 0; i < 4; i++)
#pragma unroll
			for (int j = 0; j < 4; j++) {
				int ls = scaling[i * 4 + j] * norm4(m, i, j);
				d[i * 4 + j] = (int)(((unsigned)(lv[i * 4 + j] * ls) << sh) + 8u) >> 4;
			}
	} else {
#pragma unroll
		for (int i = 0; i < 16; i++) d[i] = 0;
	}
	if (dc_override) d[0] = dc;
	int f[16];
#pragma unroll
	for (int i = 0; i < 4; i++) {
		int e0 = d[i * 4] + d[i * 4 + 2], e1 = d[i * 4] - d[i * 4 + 2];
		int e2 = (d[i * 4 + 1] >> 1) - d[i * 4 + 3], e3 = d[i * 4 + 1] + (d[i * 4 + 3] >> 1);
		f[i * 4] = e0 + e3; f[i * 4 + 1] = e1 + e2; f[i * 4 + 2] = e1 - e2; f[i * 4 + 3] = e0 - e3;
	}
#pragma unroll
	for (int j = 0; j < 4; j++) {
		int g0 = f[j] + f[8 + j], g1 = f[j] - f[8 + j];
		int g2 = (f[4 + j] >> 1) - f[12 + j], g3 = f[4 + j] + (f[12 + j] >> 1);
		dst[j] = (int16_t)sat16((g0 + g3 + 32) >> 6);
		dst[dstride + j] = (int16_t)sat16((g1 + g2 + 32) >> 6);
		dst[2 * dstride + j] = (int16_t)sat16((g1 - g2 + 32) >> 6);
		dst[3 * dstride + j] = (int16_t)sat16((g0 - g3 + 32) >> 6);
	}
}

/* one 8-point pass in the reference's int16 arithmetic (edge264_residual.c:250-296) */
__device__ __forceinline__ void idct8_1d(short a[8]) {
	short e0 = (short)(a[0] + a[4]), e1 = (short)(a[5] - a[3] - (short)((a[7] >> 1) + a[7])), e2 = (short)(a[0] - a[4]);
	short e3 = (short)(a[1] + a[7] - (short)((a[3] >> 1) + a[3])), e4 = (short)((a[2] >> 1) - a[6]);
	short e5 = (short)(a[7] - a[1] + (short)((a[5] >> 1) + a[5])), e6 = (short)((a[6] >> 1) + a[2]);
	short e7 = (short)(a[3] + a[5] + (short)((a[1] >> 1) + a[1]));
	short f0 = (short)(e0 + e6), f1 = (short)((e7 >> 2) + e1), f2 = (short)(e2 + e4), f3 = (short)((e5 >> 2) + e3);
	short f4 = (short)(e2 - e4), f5 = (short)((e3 >> 2) - e5), f6 = (short)(e0 - e6), f7 = (short)(e7 - (e1 >> 2));
	a[0] = (short)(f0 + f7); a[1] = (short)(f2 + f5); a[2] = (short)(f4 + f3); a[3] = (short)(f6 + f1);
	a[4] = (short)(f6 - f1); a[5] = (short)(f4 - f3); a[6] = (short)(f2 - f5); a[7] = (short)(f0 - f7);
}

/* cf: this macroblock's coefficient run, staged in shared memory by the caller (cp.async.bulk) */
__device__ __noinline__ void residual_stage(WarpSmem *ws, const E264MbRec *r, const E264SliceRec *sr, const int16_t *cf, int lane) {
	/* clear */
	uint4 z = make_uint4(0, 0, 0, 0);
	((uint4 *)ws->res)[lane] = z;
	if (lane < 16) ((uint4 *)ws->res)[32 + lane] = z;
	const unsigned coded = r->coded;
	const int inter = r->kind == MBK_INTER, i16 = r->kind == MBK_I16x16;
	const int qpy = r->qp[0];
	/* --- DC transforms first --- */
	if (coded & CODED_Y_DC) {
		if (lane < 16) {
			int i = lane >> 2, j = lane & 3, u = 0;
			const unsigned neg[4] = {0x0, 0xC, 0x6, 0xA};
#pragma unroll
			for (int k = 0; k < 4; k++)
#pragma unroll
				for (int l = 0; l < 4; l++) {
					int v = cf[k * 4 + l];
					int s = ((neg[i] >> k) ^ (neg[j] >> l)) & 1;
					u += s ? -v : v;
				}
			int ls = (sr->scaling4x4[0][0] * h264_norm4x4[qpy % 6][0]) << (qpy / 6);
			ws->dc[lane] = (u * ls + 32) >> 6;
		}
		cf += 16;
	} else if (lane < 16) ws->dc[lane] = 0;
	const int n_luma = (r->flags & MBF_T8x8) ? 64 * __popc(coded & 0x1111) : 16 * __popc(coded & 0xffff);
	const int16_t *cf_luma = cf;
	cf += n_luma;
	const bool any_cdc = (coded & (CODED_CB_DC | CODED_CR_DC)) != 0;
	if (lane >= 16 && lane < 24) {
		int pl = (lane - 16) >> 2, i = lane & 3, v = 0;
		if (any_cdc) {
			int a = cf[4 * pl], b = cf[4 * pl + 1], c = cf[4 * pl + 2], d = cf[4 * pl + 3];
			int f = i == 0 ? a + b + c + d : i == 1 ? a - b + c - d : i == 2 ? a + b - c - d : a - b - c + d;
			int qpc = r->qp[1 + pl];
			int ls = (sr->scaling4x4[1 + pl + inter * 3][0] * h264_norm4x4[qpc % 6][0]) << (qpc / 6);
			v = (f * ls) >> 5;
		}
		ws->dc[16 + pl * 4 + i] = v;
	}
	if (any_cdc) cf += 8;
	__syncwarp();
	/* --- luma --- */
	if (r->flags & MBF_T8x8) {
		int blk = lane >> 3, k = lane & 7;
		bool on = (coded >> (blk * 4)) & 1;
		short a[8];
		if (on) {
			int idx = __popc(coded & 0x1111 & ((1u << (blk * 4)) - 1));
			uint4 q = *(const uint4 *)(cf_luma + idx * 64 + k * 8);
			short lv[8]; *(uint4 *)lv = q;
			int div = qpy / 6, m = qpy - div * 6;
			const uint8_t *sc = sr->scaling8x8[inter] + k * 8;
#pragma unroll
			for (int j = 0; j < 8; j++) {
				int ls = sc[j] * norm8(m, k, j);
				a[j] = div < 6 ? (short)sat16((lv[j] * ls + (1 << (5 - div))) >> (6 - div)) : (short)(lv[j] * (short)(ls << (div - 6)));
			}
			idct8_1d(a);
#pragma unroll
			for (int j = 0; j < 8; j++) ws->u.t8[blk * 64 + k * 8 + j] = a[j];
		}
		__syncwarp();
		if (on) {
#pragma unroll
			for (int j = 0; j < 8; j++) a[j] = ws->u.t8[blk * 64 + j * 8 + k];
			a[0] = (short)(a[0] + 32);
			idct8_1d(a);
			int x = (blk & 1) * 8 + k, y0 = (blk >> 1) * 8;
#pragma unroll
			for (int j = 0; j < 8; j++) ws->res[(y0 + j) * 16 + x] = (short)(a[j] >> 6);
		}
		__syncwarp();
	} else if (lane < 16) {
		int b = lane;
		bool on = (coded >> b) & 1, dcov = i16 && (coded & CODED_Y_DC);
		if (on || dcov) {
			int idx = __popc(coded & ((1u << b) - 1));
			int bx = blk_x(b), by = blk_y(b);
			idct4x4(cf_luma + idx * 16, on, sr->scaling4x4[inter * 3], qpy, i16, ws->dc[by * 4 + bx], ws->res + by * 4 * 16 + bx * 4, 16);
		}
	}
	/* --- chroma AC (+DC) --- */
	if (lane >= 16 && lane < 24) {
		int j = lane - 16, pl = j >> 2, i = j & 3;
		bool on = (coded >> (16 + j)) & 1;
		if (on || any_cdc) {
			int idx = __popc((coded >> 16) & ((1u << j) - 1) & 0xff);
			idct4x4(cf + idx * 16, on, sr->scaling4x4[1 + pl + inter * 3], r->qp[1 + pl], true, ws->dc[16 + pl * 4 + i], ws->res + 256 + pl * 64 + (i >> 1) * 4 * 8 + (i & 1) * 4, 8);
		}
	}
	__syncwarp();
}

/* ------------------------------------------------------------------------------------------ */
/* intra prediction                                                                             */
/* ------------------------------------------------------------------------------------------ */
#define YT(x, y) ws->ytile[((y) + 1) * YT_STRIDE + 16 + (x)]     /* x in -1..23, y in -1..15 */
#define CT(pl, x, y) ws->ctile[pl][((y) + 1) * CT_STRIDE + 8 + (x)]

/* predicted sample (x,y) of a 4x4 block whose top-left tile coordinate is (X0,Y0) */
__device__ __noinline__ int pred4x4(const WarpSmem *ws, int X0, int Y0, int imode, int x, int y) {
	int mode = imode & 15, un = imode >> 4;
	bool hasA = !(un & 1), hasB = !(un & 2), hasC = !(un & 4);
#define T(i) ((i) < 0 ? (int)YT(X0 - 1, Y0 - 1) : (int)YT(X0 + (((i) > 3 && !hasC) ? 3 : (i)), Y0 - 1))
#define L(i) ((i) < 0 ? (int)YT(X0 - 1, Y0 - 1) : (int)YT(X0 - 1, Y0 + (i)))
	switch (mode) {
	case 0: return T(x);
	case 1: return L(y);
	case 2:
		if (hasA && hasB) return (T(0) + T(1) + T(2) + T(3) + L(0) + L(1) + L(2) + L(3) + 4) >> 3;
		if (hasA) return (L(0) + L(1) + L(2) + L(3) + 2) >> 2;
		if (hasB) return (T(0) + T(1) + T(2) + T(3) + 2) >> 2;
		return 128;
	case 3: return (x == 3 && y == 3) ? (T(6) + 3 * T(7) + 2) >> 2 : (T(x + y) + 2 * T(x + y + 1) + T(x + y + 2) + 2) >> 2;
	case 4:
		if (x > y) return (T(x - y - 2) + 2 * T(x - y - 1) + T(x - y) + 2) >> 2;
		if (x < y) return (L(y - x - 2) + 2 * L(y - x - 1) + L(y - x) + 2) >> 2;
		return (T(0) + 2 * T(-1) + L(0) + 2) >> 2;
	case 5: {
		int z = 2 * x - y;
		if (z >= 0 && !(z & 1)) return (T(x - (y >> 1) - 1) + T(x - (y >> 1)) + 1) >> 1;
		if (z > 0) return (T(x - (y >> 1) - 2) + 2 * T(x - (y >> 1) - 1) + T(x - (y >> 1)) + 2) >> 2;
		if (z == -1) return (L(0) + 2 * L(-1) + T(0) + 2) >> 2;
		return (L(y - 1) + 2 * L(y - 2) + L(y - 3) + 2) >> 2; }
	case 6: {
		int z = 2 * y - x;
		if (z >= 0 && !(z & 1)) return (L(y - (x >> 1) - 1) + L(y - (x >> 1)) + 1) >> 1;
		if (z > 0) return (L(y - (x >> 1) - 2) + 2 * L(y - (x >> 1) - 1) + L(y - (x >> 1)) + 2) >> 2;
		if (z == -1) return (L(0) + 2 * L(-1) + T(0) + 2) >> 2;
		return (T(x - 1) + 2 * T(x - 2) + T(x - 3) + 2) >> 2; }
	case 7:
		if (!(y & 1)) return (T(x + (y >> 1)) + T(x + (y >> 1) + 1) + 1) >> 1;
		return (T(x + (y >> 1)) + 2 * T(x + (y >> 1) + 1) + T(x + (y >> 1) + 2) + 2) >> 2;
	default: {
		int z = x + 2 * y;
		if (z > 5) return L(3);
		if (z == 5) return (L(2) + 3 * L(3) + 2) >> 2;
		if (!(z & 1)) return (L(y + (x >> 1)) + L(y + (x >> 1) + 1) + 1) >> 1;
		return (L(y + (x >> 1)) + 2 * L(y + (x >> 1) + 1) + L(y + (x >> 1) + 2) + 2) >> 2; }
	}
#undef T
#undef L
}

/* Intra 8x8: filtered reference samples in ws->u.edge[0][1+x] (top, x=-1..15) and edge[1][1+y] (left, y=-1..7) */
__device__ __noinline__ void intra8x8_edges(WarpSmem *ws, int X0, int Y0, int un, int lane) {
	bool hasA = !(un & 1), hasB = !(un & 2), hasC = !(un & 4), hasD = !(un & 8);
#define RT(i) ((i) < 0 ? (int)YT(X0 - 1, Y0 - 1) : (int)YT(X0 + (((i) > 7 && !hasC) ? 7 : (i)), Y0 - 1))
#define RL(i) ((i) < 0 ? (int)YT(X0 - 1, Y0 - 1) : (int)YT(X0 - 1, Y0 + (i)))
	if (lane < 16) {   /* top x = lane */
		int x = lane, v = 128;
		if (hasB) {
			if (x == 0) v = hasD ? (RT(-1) + 2 * RT(0) + RT(1) + 2) >> 2 : (3 * RT(0) + RT(1) + 2) >> 2;
			else if (x == 15) v = (RT(14) + 3 * RT(15) + 2) >> 2;
			else v = (RT(x - 1) + 2 * RT(x) + RT(x + 1) + 2) >> 2;
		}
		ws->u.edge[0][1 + x] = v;
	} else if (lane < 24) {   /* left y = lane - 16 */
		int y = lane - 16, v = 128;
		if (hasA) {
			if (y == 0) v = hasD ? (RL(-1) + 2 * RL(0) + RL(1) + 2) >> 2 : (3 * RL(0) + RL(1) + 2) >> 2;
			else if (y == 7) v = (RL(6) + 3 * RL(7) + 2) >> 2;
			else v = (RL(y - 1) + 2 * RL(y) + RL(y + 1) + 2) >> 2;
		}
		ws->u.edge[1][1 + y] = v;
	} else if (lane == 24) {
		int v = hasD ? RT(-1) : 128;
		if (hasD) {
			if (hasA && hasB) v = (RT(0) + 2 * RT(-1) + RL(0) + 2) >> 2;
			else if (hasB) v = (3 * RT(-1) + RT(0) + 2) >> 2;
			else if (hasA) v = (3 * RT(-1) + RL(0) + 2) >> 2;
		}
		ws->u.edge[0][0] = v; ws->u.edge[1][0] = v;
	}
#undef RT
#undef RL
	__syncwarp();
}
__device__ __noinline__ int pred8x8(const WarpSmem *ws, int imode, int x, int y) {
	int mode = imode & 15, un = imode >> 4;
	bool hasA = !(un & 1), hasB = !(un & 2);
#define T(i) ws->u.edge[0][1 + (i)]
#define L(i) ws->u.edge[1][1 + (i)]
	switch (mode) {
	case 0: return T(x);
	case 1: return L(y);
	case 2: {
		int st = 0, sl = 0;
#pragma unroll
		for (int k = 0; k < 8; k++) { st += T(k); sl += L(k); }
		return (hasA && hasB) ? (st + sl + 8) >> 4 : hasA ? (sl + 4) >> 3 : hasB ? (st + 4) >> 3 : 128; }
	case 3: return (x == 7 && y == 7) ? (T(14) + 3 * T(15) + 2) >> 2 : (T(x + y) + 2 * T(x + y + 1) + T(x + y + 2) + 2) >> 2;
	case 4:
		if (x > y) return (T(x - y - 2) + 2 * T(x - y - 1) + T(x - y) + 2) >> 2;
		if (x < y) return (L(y - x - 2) + 2 * L(y - x - 1) + L(y - x) + 2) >> 2;
		return (T(0) + 2 * T(-1) + L(0) + 2) >> 2;
	case 5: {
		int z = 2 * x - y;
		if (z >= 0 && !(z & 1)) return (T(x - (y >> 1) - 1) + T(x - (y >> 1)) + 1) >> 1;
		if (z > 0) return (T(x - (y >> 1) - 2) + 2 * T(x - (y >> 1) - 1) + T(x - (y >> 1)) + 2) >> 2;
		if (z == -1) return (L(0) + 2 * L(-1) + T(0) + 2) >> 2;
		return (L(y - 2 * x - 1) + 2 * L(y - 2 * x - 2) + L(y - 2 * x - 3) + 2) >> 2; }
	case 6: {
		int z = 2 * y - x;
		if (z >= 0 && !(z & 1)) return (L(y - (x >> 1) - 1) + L(y - (x >> 1)) + 1) >> 1;
		if (z > 0) return (L(y - (x >> 1) - 2) + 2 * L(y - (x >> 1) - 1) + L(y - (x >> 1)) + 2) >> 2;
		if (z == -1) return (L(0) + 2 * L(-1) + T(0) + 2) >> 2;
		return (T(x - 2 * y - 1) + 2 * T(x - 2 * y - 2) + T(x - 2 * y - 3) + 2) >> 2; }
	case 7:
		if (!(y & 1)) return (T(x + (y >> 1)) + T(x + (y >> 1) + 1) + 1) >> 1;
		return (T(x + (y >> 1)) + 2 * T(x + (y >> 1) + 1) + T(x + (y >> 1) + 2) + 2) >> 2;
	default: {
		int z = x + 2 * y;
		if (z > 13) return L(7);
		if (z == 13) return (L(6) + 3 * L(7) + 2) >> 2;
		if (!(z & 1)) return (L(y + (x >> 1)) + L(y + (x >> 1) + 1) + 1) >> 1;
		return (L(y + (x >> 1)) + 2 * L(y + (x >> 1) + 1) + L(y + (x >> 1) + 2) + 2) >> 2; }
	}
#undef T
#undef L
}

__device__ __noinline__ void intra_luma(WarpSmem *ws, const E264MbRec *r, int lane) {
	if (r->kind == MBK_I4x4) {
#pragma unroll 1
		for (int b = 0; b < 16; b++) {
			int X0 = blk_x(b) * 4, Y0 = blk_y(b) * 4;
			if (lane < 16) {
				int x = lane & 3, y = lane >> 2;
				int v = pred4x4(ws, X0, Y0, r->modes[b], x, y);
				v = clip255((short)(v + ws->res[(Y0 + y) * 16 + X0 + x]));
				YT(X0 + x, Y0 + y) = (uint8_t)v;   /* reads above touch only samples outside the block */
			}
			__syncwarp();
		}
	} else if (r->kind == MBK_I8x8) {
#pragma unroll 1
		for (int i = 0; i < 4; i++) {
			int X0 = (i & 1) * 8, Y0 = (i >> 1) * 8, im = r->modes[i];
			intra8x8_edges(ws, X0, Y0, im >> 4, lane);
			int x = lane & 7, y = lane >> 3;
			int v0 = pred8x8(ws, im, x, y), v1 = pred8x8(ws, im, x, y + 4);
			v0 = clip255((short)(v0 + ws->res[(Y0 + y) * 16 + X0 + x]));
			v1 = clip255((short)(v1 + ws->res[(Y0 + y + 4) * 16 + X0 + x]));
			__syncwarp();
			YT(X0 + x, Y0 + y) = (uint8_t)v0; YT(X0 + x, Y0 + y + 4) = (uint8_t)v1;
			__syncwarp();
		}
	} else {   /* Intra16x16 */
		int mode = r->i16_mode & 15, un = r->i16_mode >> 4;
		bool hasA = !(un & 1), hasB = !(un & 2);
		int a = 0, b = 0, c = 0, dc = 128;
		if (mode == 3) {
			int H = 0, V = 0;
#pragma unroll
			for (int k = 0; k < 8; k++) { H += (k + 1) * ((int)YT(8 + k, -1) - (int)YT(6 - k, -1)); V += (k + 1) * ((int)YT(-1, 8 + k) - (int)YT(-1, 6 - k)); }
			a = 16 * ((int)YT(-1, 15) + (int)YT(15, -1)); b = (5 * H + 32) >> 6; c = (5 * V + 32) >> 6;
		} else if (mode == 2) {
			int st = 0, sl = 0;
#pragma unroll
			for (int k = 0; k < 16; k++) { st += YT(k, -1); sl += YT(-1, k); }
			dc = (hasA && hasB) ? (st + sl + 16) >> 5 : hasA ? (sl + 8) >> 4 : hasB ? (st + 8) >> 4 : 128;
		}
#pragma unroll 1
		for (int k = 0; k < 8; k++) {   /* reads touch only samples outside the macroblock: no hazard with the writes */
			int p = lane + 32 * k, x = p & 15, y = p >> 4;
			int v = mode == 0 ? (hasB ? (int)YT(x, -1) : 128) : mode == 1 ? (hasA ? (int)YT(-1, y) : 128) : mode == 2 ? dc : clip255((a + b * (x - 7) + c * (y - 7) + 16) >> 5);
			YT(x, y) = (uint8_t)clip255((short)(v + ws->res[p]));
		}
		__syncwarp();
	}
}

__device__ __noinline__ void intra_chroma(WarpSmem *ws, const E264MbRec *r, int lane) {
	int mode = r->chroma_mode & 15, un = r->chroma_mode >> 4;
	bool hasA = !(un & 1), hasB = !(un & 2);
#pragma unroll 1
	for (int k = 0; k < 4; k++) {   /* reads touch only the row above / column left of the block */
		int p = lane + 32 * k, pl = p >> 6, x = p & 7, y = (p >> 3) & 7, v;
#define TC(i) (hasB ? (int)CT(pl, i, -1) : 128)
#define LC(i) (hasA ? (int)CT(pl, -1, i) : 128)
		if (mode == 0) {
			int bx = x >> 2, by = y >> 2, st = 0, sl = 0;
#pragma unroll
			for (int q = 0; q < 4; q++) { st += TC(bx * 4 + q); sl += LC(by * 4 + q); }
			if (bx == by) v = (hasA && hasB) ? (st + sl + 4) >> 3 : hasA ? (sl + 2) >> 2 : hasB ? (st + 2) >> 2 : 128;
			else if (bx == 1) v = hasB ? (st + 2) >> 2 : hasA ? (sl + 2) >> 2 : 128;
			else v = hasA ? (sl + 2) >> 2 : hasB ? (st + 2) >> 2 : 128;
		} else if (mode == 1) v = LC(y);
		else if (mode == 2) v = TC(x);
		else {
			int H = 0, V = 0, corner = CT(pl, -1, -1);
#pragma unroll
			for (int q = 0; q < 4; q++) {
				H += (q + 1) * ((int)CT(pl, 4 + q, -1) - (q == 3 ? corner : (int)CT(pl, 2 - q, -1)));
				V += (q + 1) * ((int)CT(pl, -1, 4 + q) - (q == 3 ? corner : (int)CT(pl, -1, 2 - q)));
			}
			int a = 16 * ((int)CT(pl, -1, 7) + (int)CT(pl, 7, -1)), b = (34 * H + 32) >> 6, c = (34 * V + 32) >> 6;
			v = clip255((a + b * (x - 3) + c * (y - 3) + 16) >> 5);
		}
#undef TC
#undef LC
		CT(pl, x, y) = (uint8_t)clip255((short)(v + ws->res[256 + pl * 64 + y * 8 + x]));
	}
	__syncwarp();
}

/* ------------------------------------------------------------------------------------------ */
/* inter prediction                                                                             */
/* ------------------------------------------------------------------------------------------ */
__device__ __forceinline__ int tap6(int a, int b, int c, int d, int e, int f) { return a - 5 * b + 20 * c + 20 * d - 5 * e + f; }

/* Motion compensation of one macroblock = a list of square rectangles (16, 8 or 4 luma samples wide, one
 * motion vector and one reference each).  Each rectangle needs a (S+5)^2 luma window and two (S/2+1)^2 chroma
 * windows of its reference picture.  Windows that lie inside the picture are fetched by the TMA unit
 * (cp.async.bulk.tensor.3d over (x, y, frame slot), one luma and two chroma boxes counted on one mbarrier);
 * windows that touch the border are gathered sample by sample with clamped coordinates (8.4.2.2.1).  Two window
 * buffers per warp: the fetch of rectangle i+1 is in flight while rectangle i is filtered. */
struct McCtx {
	uint8_t *win0;                 /* two window buffers, WIN_BYTES apart */
	unsigned long long *bar0;      /* two mbarriers, adjacent */
	unsigned parity;               /* bit b: phase parity of barrier b */
	unsigned pending;              /* bit b: buffer b is being filled by the TMA unit */
	unsigned offs[2];              /* per buffer: lo | cob << 8 | cor << 16 — where the wanted luma / Cb / Cr window starts inside its rows */
};
/* one box of frame slot z (third tensor coordinate); x must be a multiple of 16 (bytes) */
__device__ __forceinline__ void tma_load_box(void *dst, const void *tmap, int x, int y, int z, void *bar) {
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
	             :: "r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

/* rectangle code: bit0 list, bits1-2 size class (0:16 1:8 2:4), bits3-4 x0/4, bits5-6 y0/4 */
#define RECT(l, sc, x0, y0) ((l) | ((sc) << 1) | (((x0) >> 2) << 3) | (((y0) >> 2) << 5))

__device__ __forceinline__ void mc_issue(WarpSmem *ws, McCtx &mc, int b, const PicJob &J, const E264MbRec *r, int mbx, int mby, int rect, int lane) {
	const int l = rect & 1, sc = (rect >> 1) & 3, x0 = ((rect >> 3) & 3) << 2, y0 = ((rect >> 5) & 3) << 2, S = 16 >> sc;
	const int WW = S + 5, CWW = (S >> 1) + 1;
	const int z0 = blk_z(x0 >> 2, y0 >> 2);
	const int mvx = r->mv[l][z0][0], mvy = r->mv[l][z0][1];
	int slot = r->ref_pic[l][z0 >> 2];
	if (slot < 0 || slot >= J.n_slots) slot = J.dst_slot;
	const int W = J.w_mbs * 16, H = J.h_mbs * 16;
	const int X0 = mbx * 16 + x0 + (mvx >> 2) - 2, Y0 = mby * 16 + y0 + (mvy >> 2) - 2;
	const int CX0 = mbx * 8 + (x0 >> 1) + (mvx >> 3), CY0 = mby * 8 + (y0 >> 1) + (mvy >> 3);
	uint8_t *win = mc.win0 + b * WIN_BYTES;
	unsigned long long *bar = mc.bar0 + b;
	const bool interior = J.tmaps != nullptr && X0 >= 0 && Y0 >= 0 && X0 + WW <= W && Y0 + WW <= H && CX0 >= 0 && CY0 >= 0 && CX0 + CWW <= (W >> 1) && CY0 + CWW <= (H >> 1);
	mc.pending = (mc.pending & ~(1u << b)) | ((unsigned)interior << b);
	unsigned offs = 0;
	if (interior) {
		const int crx = CX0 + (J.stride_c >> 1);
		offs = (unsigned)(X0 & 15) | (unsigned)(CX0 & 15) << 8 | (unsigned)(crx & 15) << 16;
		if (lane == 0) {
			const char *tm = (const char *)J.tmaps + sc * 128;
			const unsigned bytes = (unsigned)(WIN_STRIDE * WW + 2 * WIN_C_STRIDE * CWW);
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
			tma_load_box(win, tm, X0 & ~15, Y0, slot, bar);
			tma_load_box(win + WIN_CB_OFF, tm + 3 * 128, CX0 & ~15, CY0, slot, bar);
			tma_load_box(win + WIN_CR_OFF, tm + 3 * 128, crx & ~15, CY0, slot, bar);
		}
	} else {
		const uint8_t *ref = J.frames + (size_t)slot * J.frame_bytes;
		const int NL = WW * WW, NC1 = CWW * CWW;
		const int rw = (65536 + WW - 1) / WW, rcw = (65536 + CWW - 1) / CWW;   /* exact reciprocals for i < 512 */
#pragma unroll 1
		for (int i = lane; i < NL; i += 32) {
			int row = (i * rw) >> 16, col = i - row * WW;
			int xx = min(max(X0 + col, 0), W - 1), yy = min(max(Y0 + row, 0), H - 1);
			win[row * WIN_STRIDE + col] = __ldg(ref + (size_t)yy * J.stride_y + xx);
		}
#pragma unroll 1
		for (int i = lane; i < 2 * NC1; i += 32) {
			int pl = i >= NC1, j = i - pl * NC1, row = (j * rcw) >> 16, col = j - row * CWW;
			int xx = min(max(CX0 + col, 0), (W >> 1) - 1), yy = min(max(CY0 + row, 0), (H >> 1) - 1);
			win[(pl ? WIN_CR_OFF : WIN_CB_OFF) + row * WIN_C_STRIDE + col] = __ldg(ref + J.plane_y + pl * (J.stride_c >> 1) + (size_t)yy * J.stride_c + xx);
		}
	}
	if (b) mc.offs[1] = offs; else mc.offs[0] = offs;
}

/* filter the rectangle whose windows sit in buffer b; the plain prediction goes to the macroblock tile, or to
 * ws->pt1 when it is the list-1 half of a bi-predicted quadrant (mc_blend combines them).  false = the TMA never
 * completed. */
__device__ __forceinline__ bool mc_compute(WarpSmem *ws, McCtx &mc, int b, const E264MbRec *r, int rect, int lane PH_ARGS) {
	const int l = rect & 1, sc = (rect >> 1) & 3, x0 = ((rect >> 3) & 3) << 2, y0 = ((rect >> 5) & 3) << 2, S = 16 >> sc;
	const int CW = S >> 1;
	const int z0 = blk_z(x0 >> 2, y0 >> 2);
	const int mvx = r->mv[l][z0][0], mvy = r->mv[l][z0][1];
	const uint8_t *win = mc.win0 + b * WIN_BYTES;
	const unsigned offs = b ? mc.offs[1] : mc.offs[0];
	if ((mc.pending >> b) & 1) {
		if (!mbar_wait(mc.bar0 + b, (mc.parity >> b) & 1)) return false;
		mc.parity ^= 1u << b; mc.pending &= ~(1u << b);
	} else __syncwarp();
	PH(4);
	const int i8r = ((y0 >> 3) << 1) | (x0 >> 3);
	const bool second = l == 1 && r->ref_idx[0][i8r] >= 0;
	uint8_t *dl = second ? ws->pt1 + y0 * 16 + x0 : &YT(x0, y0);
	const int dls = second ? 16 : YT_STRIDE;
	const int fx = mvx & 3, fy = mvy & 3, sh = 4 - sc;
	const uint8_t *wl = win + (offs & 15) + 2 * WIN_STRIDE + 2;
	/* one loop per class of fractional position: the branch is uniform for the rectangle */
#define HSUM(g) tap6((g)[-2], (g)[-1], (g)[0], (g)[1], (g)[2], (g)[3])
#define VSUM(g) tap6((g)[-2 * WIN_STRIDE], (g)[-WIN_STRIDE], (g)[0], (g)[WIN_STRIDE], (g)[2 * WIN_STRIDE], (g)[3 * WIN_STRIDE])
#define LUMA_LOOP(EXPR) _Pragma("unroll 1") for (int p = lane; p < S * S; p += 32) { int x = p & (S - 1), y = p >> sh; const uint8_t *g = wl + y * WIN_STRIDE + x; dl[y * dls + x] = (uint8_t)(EXPR); }
	if (!(fx | fy)) { LUMA_LOOP(g[0]) }
	else if (!fy) {   /* a, b, c */
		const int o2 = fx == 3;
		if (fx == 2) { LUMA_LOOP(clip255((HSUM(g) + 16) >> 5)) } else { LUMA_LOOP((clip255((HSUM(g) + 16) >> 5) + g[o2] + 1) >> 1) }
	} else if (!fx) {   /* d, h, n */
		const int o2 = fy == 3 ? WIN_STRIDE : 0;
		if (fy == 2) { LUMA_LOOP(clip255((VSUM(g) + 16) >> 5)) } else { LUMA_LOOP((clip255((VSUM(g) + 16) >> 5) + g[o2] + 1) >> 1) }
	} else if ((fx & 1) && (fy & 1)) {   /* e, g, p, r: horizontal half of row y or y+1, vertical half of column x or x+1 */
		const int ro = fy == 3 ? WIN_STRIDE : 0, cofs = fx == 3;
		LUMA_LOOP((clip255((HSUM(g + ro) + 16) >> 5) + clip255((VSUM(g + cofs) + 16) >> 5) + 1) >> 1)
	} else {   /* f, i, j, k, q: centre sample, combined like the reference (int16 wrap, edge264_inter.c:4-9) */
		const bool vfirst = fx & 1;
		const int sa = vfirst ? 1 : WIN_STRIDE, sb = vfirst ? WIN_STRIDE : 1;
		const int snd = fx == 2 ? (fy == 2 ? 0 : 1) : 2;          /* 0 none (j), 1 horizontal half b/s, 2 vertical half h/m */
		const int so = fx == 2 ? (fy == 3 ? WIN_STRIDE : 0) : (fx == 3 ? 1 : 0);
#pragma unroll 1
		for (int p = lane; p < S * S; p += 32) {
			int x = p & (S - 1), y = p >> sh; const uint8_t *g = wl + y * WIN_STRIDE + x;
			int t[6];
#pragma unroll
			for (int k = 0; k < 6; k++) { const uint8_t *q = g + (k - 2) * sa; t[k] = tap6(q[-2 * sb], q[-sb], q[0], q[sb], q[2 * sb], q[3 * sb]); }
			int af = t[0] + t[5], be = t[1] + t[4], cd = t[2] + t[3];
			int t16 = (short)(((af - be) >> 2) + (cd - be));
			int v = clip255(((t16 >> 2) + cd + 32) >> 6);
			if (snd == 1) v = (v + clip255((HSUM(g + so) + 16) >> 5) + 1) >> 1;
			else if (snd == 2) v = (v + clip255((VSUM(g + so) + 16) >> 5) + 1) >> 1;
			dl[y * dls + x] = (uint8_t)v;
		}
	}
	__syncwarp(); PH(5);
	const int cfx = mvx & 7, cfy = mvy & 7;
	const int cA = (8 - cfx) * (8 - cfy), cB = cfx * (8 - cfy), cC = (8 - cfx) * cfy, cD = cfx * cfy;
	const int cx0 = x0 >> 1, cy0 = y0 >> 1;
#pragma unroll 1
	for (int p = lane; p < 2 * CW * CW; p += 32) {
		int pl = p >= CW * CW, q = p - pl * CW * CW, x = q & (CW - 1), y = q >> (sh - 1);
		const uint8_t *cwn = win + (pl ? WIN_CR_OFF + ((offs >> 16) & 15) : WIN_CB_OFF + ((offs >> 8) & 15)) + y * WIN_C_STRIDE + x;
		int v = (cA * cwn[0] + cB * cwn[1] + cC * cwn[WIN_C_STRIDE] + cD * cwn[WIN_C_STRIDE + 1] + 32) >> 6;
		if (second) ws->pt1[256 + pl * 64 + (cy0 + y) * 8 + cx0 + x] = (uint8_t)v; else CT(pl, cx0 + x, cy0 + y) = (uint8_t)v;
	}
#undef LUMA_LOOP
#undef HSUM
#undef VSUM
	__syncwarp();
	PH(6);
	return true;
}

/* 8.4.2.3 weighted sample prediction, once per macroblock after all rectangles are filtered.  Per 8x8 quadrant
 * and component, mode: 0 keep, 1 default average of the two predictions, 2 explicit weight on the single
 * prediction, 3 weighted sum of both (explicit, or implicit with weights 64-w1, w1, shift 6).  Same integer
 * formulas as the reference's five schemes (edge264_inter.c:1140-1197) after its offset folding. */
__device__ __forceinline__ void mc_blend(WarpSmem *ws, const E264MbRec *r, const E264SliceRec *sr, int lane) {
	bool any = false;
	if (lane < 12) {
		const int i8 = lane & 3, c = lane >> 2;
		const int r0 = r->ref_idx[0][i8], r1 = r->ref_idx[1][i8];
		int mode = 0, w0 = 0, w1 = 0, o = 0, lw = 0;
		if (r0 >= 0 && r1 >= 0) {
			if (sr->wp_mode == WP_EXPLICIT) {
				mode = 3; w0 = sr->wp_w[0][r0 & 15][c]; w1 = sr->wp_w[1][r1 & 15][c];
				o = (sr->wp_o[0][r0 & 15][c] + sr->wp_o[1][r1 & 15][c] + 1) >> 1; lw = c ? sr->chroma_log2_wd : sr->luma_log2_wd;
			} else if (sr->wp_mode == WP_IMPLICIT) {
				mode = 3; w1 = sr->implicit_w1[r0 & 15][r1 & 15]; w0 = 64 - w1; lw = 5;
			} else mode = 1;
		} else if (sr->wp_mode == WP_EXPLICIT && (r0 >= 0 || r1 >= 0)) {
			const int l = r0 >= 0 ? 0 : 1, ri = (l ? r1 : r0) & 15;
			mode = 2; w1 = sr->wp_w[l][ri][c]; o = sr->wp_o[l][ri][c]; lw = c ? sr->chroma_log2_wd : sr->luma_log2_wd;
		}
		ws->wq[i8][c][0] = mode; ws->wq[i8][c][1] = w0; ws->wq[i8][c][2] = w1; ws->wq[i8][c][3] = (o & 0xffff) | (lw << 16);
		any = mode != 0;
	}
	if (!__any_sync(0xffffffffu, any)) return;
	__syncwarp();
#pragma unroll 1
	for (int p = lane; p < 384; p += 32) {
		int c, i8, a; uint8_t *dst;
		if (p < 256) { int x = p & 15, y = p >> 4; c = 0; i8 = ((y >> 3) << 1) | (x >> 3); dst = &YT(x, y); }
		else { int q = p - 256, pl = q >> 6, x = q & 7, y = (q >> 3) & 7; c = 1 + pl; i8 = ((y >> 2) << 1) | (x >> 2); dst = &CT(pl, x, y); }
		const int b2 = ws->pt1[p];
		const int4 wv = *(const int4 *)ws->wq[i8][c];
		const int mode = wv.x, o = (short)(wv.w & 0xffff), lw = wv.w >> 16;
		a = *dst;
		if (mode == 1) a = (a + b2 + 1) >> 1;
		else if (mode == 2) a = clip255((lw >= 1 ? ((a * wv.z + (1 << (lw - 1))) >> lw) : a * wv.z) + o);
		else if (mode == 3) a = clip255(((a * wv.y + b2 * wv.z + (1 << lw)) >> (lw + 1)) + o);
		*dst = (uint8_t)a;
	}
	__syncwarp();
}

/* list the rectangles of this macroblock (list 0 first: list 1 blends with what list 0 stored) */
__device__ __forceinline__ int mc_rects(const E264MbRec *r, uint8_t *out, int lane) {
	int n = 0;
	for (int l = 0; l < 2; l++) {
		int z = lane & 15;
		bool same = r->mv[l][z][0] == r->mv[l][0][0] && r->mv[l][z][1] == r->mv[l][0][1] && r->ref_idx[l][z >> 2] == r->ref_idx[l][0] && r->ref_idx[l ^ 1][z >> 2] == r->ref_idx[l ^ 1][0];
		if (__all_sync(0xffffffffu, same)) {
			if (r->ref_idx[l][0] >= 0) { if (lane == 0) out[n] = (uint8_t)RECT(l, 0, 0, 0); n++; }
			continue;
		}
		for (int i8 = 0; i8 < 4; i8++) {
			if (r->ref_idx[l][i8] < 0) continue;
			int zb = i8 * 4, x0 = (i8 & 1) * 8, y0 = (i8 >> 1) * 8;
			bool s8 = true;
			for (int k = 1; k < 4; k++) s8 = s8 && r->mv[l][zb + k][0] == r->mv[l][zb][0] && r->mv[l][zb + k][1] == r->mv[l][zb][1];
			if (s8) { if (lane == 0) out[n] = (uint8_t)RECT(l, 1, x0, y0); n++; }
			else for (int k = 0; k < 4; k++) { if (lane == 0) out[n] = (uint8_t)RECT(l, 2, x0 + (k & 1) * 4, y0 + (k >> 1) * 4); n++; }
		}
	}
	__syncwarp();
	return n;
}

__device__ __noinline__ bool inter_predict(WarpSmem *ws, McCtx &mc, uint8_t *rects, const PicJob &J, const E264MbRec *r, const E264SliceRec *sr, int mbx, int mby, int lane PH_ARGS) {
	const int n = mc_rects(r, rects, lane);
	PH(2);
#pragma unroll 1
	for (int i = -1; i < n; i++) {
		if (i + 1 < n) { mc_issue(ws, mc, (i + 1) & 1, J, r, mbx, mby, rects[i + 1], lane); PH(3); }
		if (i >= 0 && !mc_compute(ws, mc, i & 1, r, rects[i], lane PH_PASS)) return false;
	}
	mc_blend(ws, r, sr, lane);
	/* add the residual */
#pragma unroll
	for (int k = 0; k < 8; k++) { int p = lane + 32 * k, x = p & 15, y = p >> 4; YT(x, y) = (uint8_t)clip255((short)((int)YT(x, y) + ws->res[p])); }
#pragma unroll
	for (int k = 0; k < 4; k++) { int p = lane + 32 * k, pl = p >> 6, x = p & 7, y = (p >> 3) & 7; CT(pl, x, y) = (uint8_t)clip255((short)((int)CT(pl, x, y) + ws->res[256 + p])); }
	__syncwarp();
	PH(7);
	return true;
}

/* ------------------------------------------------------------------------------------------ */
/* reconstruction kernel                                                                        */
/* ------------------------------------------------------------------------------------------ */
/* load the macroblock's residual (written by e264_residual_kernel) into ws->res, or clear it */
__device__ __forceinline__ void fetch_residual(WarpSmem *ws, const PicJob &J, const E264MbRec *r, int mb, int lane) {
	const uint4 *src = (const uint4 *)(J.resid + (size_t)mb * 384);
	const bool on = r->coded != 0;
	uint4 z = make_uint4(0, 0, 0, 0);
	((uint4 *)ws->res)[lane] = on ? __ldg(src + lane) : z;
	if (lane < 16) ((uint4 *)ws->res)[32 + lane] = on ? __ldg(src + 32 + lane) : z;
	__syncwarp();
}

__device__ __forceinline__ void store_mb(WarpSmem *ws, const PicJob &J, uint8_t *Y, uint8_t *C, int lane) {
	const int cpl = J.stride_c >> 1;
	if (lane < 16) *(uint4 *)(Y + (size_t)lane * J.stride_y) = *(const uint4 *)&YT(0, lane);
	else { int j = lane - 16, pl = j >> 3, row = j & 7; *(uint2 *)(C + pl * cpl + (size_t)row * J.stride_c) = *(const uint2 *)&CT(pl, 0, row); }
}

/* ---- kernel 1: inverse quantisation + inverse transforms of every coded macroblock (no dependencies) ----
 * Each warp owns a two-stage pipeline: while macroblock i is transformed, the coefficient run of its next
 * macroblock (16..816 bytes, contiguous in the pool, 16-byte aligned) is in flight as one TMA bulk copy
 * (cp.async.bulk global -> shared, completion counted in bytes on the warp's mbarrier). */
#define RES_COEF_MAX 416   /* 16 + 256 + 8 + 128 levels, rounded up */

/* number of int16 levels a record owns in the pool (same layout rule as sx_pool_take on the host) */
__device__ __forceinline__ int rec_coef_count(const E264MbRec *r) {
	const unsigned coded = r->coded;
	int n = (coded & CODED_Y_DC) ? 16 : 0;
	n += (r->flags & MBF_T8x8) ? 64 * __popc(coded & 0x1111) : 16 * __popc(coded & 0xffff);
	if (coded & (CODED_CB_DC | CODED_CR_DC)) n += 8;
	n += 16 * __popc((coded >> 16) & 0xff);
	return n;
}

struct __align__(16) ResStage {
	uint4 rec4[12];
	int16_t coef[RES_COEF_MAX];
};

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) e264_residual_kernel(PicJob J) {
	TraceScope trace_(J, 0);
	__shared__ WarpSmem smem[WARPS_PER_BLOCK];
	__shared__ ResStage stage[WARPS_PER_BLOCK][2];
	__shared__ __align__(8) unsigned long long bars[WARPS_PER_BLOCK][2];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	WarpSmem *ws = &smem[w];
	const int nmb = J.w_mbs * J.h_mbs, step = gridDim.x * WARPS_PER_BLOCK;
	if (lane == 0) {
		mbar_init(&bars[w][0], 1); mbar_init(&bars[w][1], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncwarp();
	/* stage s <- macroblock m: record through the read-only path, coefficient run through TMA */
	auto issue = [&](int m, int s) -> bool {
		ResStage *st = &stage[w][s];
		if (lane < 12) st->rec4[lane] = __ldg((const uint4 *)(J.recs + m) + lane);
		__syncwarp();
		const E264MbRec *r = (const E264MbRec *)st->rec4;
		const bool on = r->coded != 0 && r->kind != MBK_IPCM && (J.resid_inter || r->kind != MBK_INTER);
		if (on && lane == 0) tma_bulk_g2s(st->coef, J.coefs + r->coef_off, (unsigned)rec_coef_count(r) * 2u, &bars[w][s]);
		return on;
	};
	int mb = blockIdx.x * WARPS_PER_BLOCK + w;
	if (mb >= nmb) return;
	unsigned parity[2] = {0, 0};
	int s = 0;
	bool on = issue(mb, 0);
	for (; mb < nmb; mb += step, s ^= 1) {
		bool on_next = false;
		if (mb + step < nmb) on_next = issue(mb + step, s ^ 1);
		if (on) {
			ResStage *st = &stage[w][s];
			if (!mbar_wait(&bars[w][s], parity[s])) { if (lane == 0) atomicExch(J.err, 3u); return; }
			parity[s] ^= 1;
			const E264MbRec *r = (const E264MbRec *)st->rec4;
			residual_stage(ws, r, J.slices + r->slice_idx, st->coef, lane);
			uint4 *dst = (uint4 *)(J.resid + (size_t)mb * 384);
			dst[lane] = ((const uint4 *)ws->res)[lane];
			if (lane < 16) dst[32 + lane] = ((const uint4 *)ws->res)[32 + lane];
		}
		__syncwarp();
		on = on_next;
	}
}

/* ---- kernel 2: inter macroblocks: motion compensation + weighting + residual (no dependencies) ---- */
template <int MINB>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, MINB) e264_inter_kernel(PicJob J) {
	TraceScope trace_(J, 1);
	__shared__ WarpSmem smem[WARPS_PER_BLOCK];
	__shared__ __align__(128) uint8_t wins[WARPS_PER_BLOCK][2][WIN_BYTES];
	__shared__ __align__(8) unsigned long long bars[WARPS_PER_BLOCK][2];
	__shared__ uint8_t rects[WARPS_PER_BLOCK][32];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	WarpSmem *ws = &smem[w];
	McCtx mc;
	mc.win0 = wins[w][0]; mc.bar0 = &bars[w][0];
	mc.parity = 0; mc.pending = 0; mc.offs[0] = mc.offs[1] = 0;
	if (lane == 0) {
		mbar_init(mc.bar0, 1); mbar_init(mc.bar0 + 1, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	/* the tensor maps were written by the host (cudaMemcpy): acquire them for the tensormap proxy once per block */
	if (J.tmaps != nullptr) {
		for (int i = threadIdx.x; i < 6; i += blockDim.x) {
			asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" :: "l"((const char *)J.tmaps + (size_t)i * 128) : "memory");
		}
	}
	__syncthreads();
	const int nmb = J.w_mbs * J.h_mbs;
	uint8_t *dst = J.frames + (size_t)J.dst_slot * J.frame_bytes;
	PH_DECL
	/* tickets are drawn one macroblock ahead: the atomic's round trip overlaps the previous macroblock */
	unsigned tnext = 0;
	if (lane == 0) tnext = atomicAdd(J.tickets, 1u);
	for (;;) {
		const unsigned t = __shfl_sync(0xffffffffu, tnext, 0);
		if (t >= (unsigned)nmb) break;
		if (lane == 0) tnext = atomicAdd(J.tickets, 1u);
		const int mb = (int)t, mbx = mb % J.w_mbs, mby = mb / J.w_mbs;
		if (__ldg(&J.recs[mb].kind) != MBK_INTER) continue;
		if (lane < 12) ws->rec4[lane] = __ldg((const uint4 *)(J.recs + mb) + lane);
		__syncwarp();
		const E264MbRec *r = (const E264MbRec *)ws->rec4;
		PH(0);
		fetch_residual(ws, J, r, mb, lane);
		PH(1);
		if (!inter_predict(ws, mc, rects[w], J, r, J.slices + r->slice_idx, mbx, mby, lane PH_PASS)) { if (lane == 0) atomicExch(J.err, 4u); break; }
		store_mb(ws, J, dst + (size_t)(mby * 16) * J.stride_y + mbx * 16, dst + J.plane_y + (size_t)(mby * 8) * J.stride_c + mbx * 8, lane);
		if (lane == 0) J.flags[mb] = J.epoch;     /* visible to the intra kernel through the kernel boundary */
		__syncwarp();
		PH(8);
	}
	PH_FLUSH(J);
}

/* ---- kernel 3: intra (and I_PCM) macroblocks, in dependency order ----
 * rows_mode == 0 (few intra macroblocks): tickets over macroblocks, each intra one waits for A, D, B, C
 * (inter neighbours were flagged by kernel 2).  rows_mode == 1 (intra pictures): one warp walks a row, the left
 * neighbour's column stays in shared memory, only C (or B at the row's end) is waited for. */
__device__ __forceinline__ void intra_mb(WarpSmem *ws, const PicJob &J, uint8_t *dst, int mb, int mbx, int mby, int lane, bool rows_mode) {
	if (lane < 12) ws->rec4[lane] = __ldg((const uint4 *)(J.recs + mb) + lane);
	__syncwarp();
	const E264MbRec *r = (const E264MbRec *)ws->rec4;
	uint8_t *Y = dst + (size_t)(mby * 16) * J.stride_y + mbx * 16;
	uint8_t *C = dst + J.plane_y + (size_t)(mby * 8) * J.stride_c + mbx * 8;
	const int kind = r->kind, cpl = J.stride_c >> 1;
	if (kind == MBK_INTER) {   /* rows mode only: already reconstructed; pick its right column up for the carry */
		if (lane < 16) YT(15, lane) = __ldcg(Y + (size_t)lane * J.stride_y + 15);
		else { int j = lane - 16; CT(j >> 3, 7, j & 7) = __ldcg(C + (j >> 3) * cpl + (size_t)(j & 7) * J.stride_c + 7); }
		__syncwarp();
	} else if (kind == MBK_IPCM) {
		const uint8_t *s = (const uint8_t *)(J.coefs + r->coef_off);
		if (lane < 16) *(uint4 *)&YT(0, lane) = __ldg((const uint4 *)s + lane);
		else { int j = lane - 16; *(uint2 *)&CT(j >> 3, 0, j & 7) = __ldg((const uint2 *)(s + 256) + j); }
		__syncwarp();
		store_mb(ws, J, Y, C, lane);
	} else {
		fetch_residual(ws, J, r, mb, lane);
		bool ok = true;
		if (lane == 0) {
			if (!rows_mode) {
				if (mbx > 0) ok = wait_flag(J.flags, mb - 1, J.epoch, J.err);
				if (ok && mby > 0 && mbx > 0) ok = wait_flag(J.flags, mb - J.w_mbs - 1, J.epoch, J.err);
				if (ok && mby > 0) ok = wait_flag(J.flags, mb - J.w_mbs, J.epoch, J.err);
			}
			if (rows_mode) {
				/* the warp of the row above must have passed C (or the end of its row): its counter orders B and D before
				 * it, whatever kernel reconstructed C — the flag of an inter C says nothing about an intra B */
				if (mby > 0) ok = wait_progress(J.flags + J.w_mbs * J.h_mbs + 2 * J.h_mbs + mby - 1, J.epoch * 2048u + (unsigned)min(mbx + 2, J.w_mbs), J.err);
			} else if (ok && mby > 0) ok = wait_flag(J.flags, mbx < J.w_mbs - 1 ? mb - J.w_mbs + 1 : mb - J.w_mbs, J.epoch, J.err);
			__threadfence();
		}
		__syncwarp();
		const bool up = mby > 0, left = mbx > 0, carried = rows_mode && left;
		if (up) {
			int x = lane - 1;   /* -1..23; the corner is always fetched (the previous macroblock may not have loaded its top row) */
			if (x < 24 && (x >= 0 || left) && (x < 16 || mbx < J.w_mbs - 1)) YT(x, -1) = __ldcg(Y - J.stride_y + x);
			if (lane < 18) { int pl = lane / 9, cx = lane % 9 - 1; if (cx >= 0 || left) CT(pl, cx, -1) = __ldcg(C + pl * cpl - J.stride_c + cx); }
		}
		if (left && !carried) {
			if (lane < 16) YT(-1, lane) = __ldcg(Y + (size_t)lane * J.stride_y - 1);
			else { int j = lane - 16, pl = j >> 3, row = j & 7; CT(pl, -1, row) = __ldcg(C + pl * cpl + (size_t)row * J.stride_c - 1); }
		}
		__syncwarp();
		intra_luma(ws, r, lane);
		intra_chroma(ws, r, lane);
		store_mb(ws, J, Y, C, lane);
	}
	__syncwarp();
	if (rows_mode) {   /* right-most column becomes the next macroblock's left neighbour */
		uint8_t v = 0;
		if (lane < 16) v = YT(15, lane);
		else if (lane < 24) v = CT(0, 7, lane - 16);
		uint8_t v2 = lane < 8 ? CT(1, 7, lane) : 0;
		__syncwarp();
		if (lane < 16) YT(-1, lane) = v;
		else if (lane < 24) CT(0, -1, lane - 16) = v;
		if (lane < 8) CT(1, -1, lane) = v2;
	}
	if (lane == 0) {
		if (kind != MBK_INTER) { __threadfence(); *(volatile unsigned *)(J.flags + mb) = J.epoch; }
		if (rows_mode) *(volatile unsigned *)(J.flags + J.w_mbs * J.h_mbs + 2 * J.h_mbs + mby) = J.epoch * 2048u + (unsigned)mbx + 1u;   /* after the fence above when this macroblock wrote samples */
	}
	__syncwarp();
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) e264_intra_kernel(PicJob J) {
	TraceScope trace_(J, 2);
	__shared__ WarpSmem smem[WARPS_PER_BLOCK];
	const int lane = threadIdx.x & 31;
	WarpSmem *ws = &smem[threadIdx.x >> 5];
	const int nmb = J.w_mbs * J.h_mbs;
	uint8_t *dst = J.frames + (size_t)J.dst_slot * J.frame_bytes;
	for (;;) {
		unsigned t = 0;
		if (lane == 0) t = atomicAdd(J.tickets + 2, 1u);
		t = __shfl_sync(0xffffffffu, t, 0);
		if (J.rows_mode) {
			if (t >= (unsigned)J.h_mbs) break;
			for (int mbx = 0; mbx < J.w_mbs; mbx++) intra_mb(ws, J, dst, (int)t * J.w_mbs + mbx, mbx, (int)t, lane, true);
		} else {
			if (t >= (unsigned)J.n_intra) break;
			const int mb = (int)__ldg(J.intra_list + t);
			if (mb >= nmb) continue;
			intra_mb(ws, J, dst, mb, mb % J.w_mbs, mb / J.w_mbs, lane, false);
		}
	}
}

/* ------------------------------------------------------------------------------------------ */
/* deblocking kernel                                                                            */
/* ------------------------------------------------------------------------------------------ */
struct __align__(16) DbSmem {
	uint8_t ypix[20 * 32];        /* rows -4..15, cols -4..15 at [ (r+4)*32 + 12 + (c+4) ] -> sample (0,r) at 16-byte aligned offset 16 */
	uint8_t cpix[2][10 * 16];     /* rows -2..7, cols -2..7 at [(r+2)*16 + 6 + (c+2)] -> sample (0,r) at offset 8 */
	int8_t bs[32];             /* [dir][edge][segment] */
	uint8_t alpha[3][3], beta[3][3];   /* [plane][0 internal, 1 left edge, 2 top edge] */
	uint8_t ia[3][3];
	uint4 rq[12], rl[12], rt[12];   /* records of the current, left and top macroblocks (row-walking kernel) */
};
#define DY(x, y) ds->ypix[((y) + 4) * 32 + 16 + (x)]
#define DC_(pl, x, y) ds->cpix[pl][((y) + 2) * 16 + 8 + (x)]

__device__ __forceinline__ int iabs_(int v) { return v < 0 ? -v : v; }

__device__ int bs_pair(const E264MbRec *p, int bp, const E264MbRec *q, int bq, bool mb_edge) {
	if (p->kind != MBK_INTER || q->kind != MBK_INTER) return mb_edge ? 4 : 3;
	if (((p->coded >> bp) & 1) || ((q->coded >> bq) & 1)) return 2;
	int p0 = p->ref_idx[0][bp >> 2] < 0 ? -1 : p->ref_pic[0][bp >> 2], p1 = p->ref_idx[1][bp >> 2] < 0 ? -1 : p->ref_pic[1][bp >> 2];
	int q0 = q->ref_idx[0][bq >> 2] < 0 ? -1 : q->ref_pic[0][bq >> 2], q1 = q->ref_idx[1][bq >> 2] < 0 ? -1 : q->ref_pic[1][bq >> 2];
	if (!((p0 == q0 && p1 == q1) || (p0 == q1 && p1 == q0))) return 1;
#define FAR(lp, lq) (iabs_(p->mv[lp][bp][0] - q->mv[lq][bq][0]) >= 4 || iabs_(p->mv[lp][bp][1] - q->mv[lq][bq][1]) >= 4)
	if (p0 >= 0 && p1 >= 0) {
		if (p0 != p1) return (p0 == q0) ? (FAR(0, 0) || FAR(1, 1)) : (FAR(0, 1) || FAR(1, 0));
		return (FAR(0, 0) || FAR(1, 1)) && (FAR(0, 1) || FAR(1, 0));
	}
	int lp = p0 >= 0 ? 0 : 1, lq = q0 >= 0 ? 0 : 1;
	return FAR(lp, lq);
#undef FAR
}

__device__ __forceinline__ void filter_luma(uint8_t *pix, int step, int bs, int alpha, int beta, int tc0) {
	int p0 = pix[-step], p1 = pix[-2 * step], p2 = pix[-3 * step], q0 = pix[0], q1 = pix[step], q2 = pix[2 * step];
	if (!(iabs_(p0 - q0) < alpha && iabs_(p1 - p0) < beta && iabs_(q1 - q0) < beta)) return;
	int ap = iabs_(p2 - p0), aq = iabs_(q2 - q0);
	if (bs < 4) {
		int tc = tc0 + (ap < beta) + (aq < beta);
		int d = min(max((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc), tc);
		pix[-step] = (uint8_t)clip255(p0 + d); pix[0] = (uint8_t)clip255(q0 - d);
		if (ap < beta) pix[-2 * step] = (uint8_t)(p1 + min(max((p2 + ((p0 + q0 + 1) >> 1) - (p1 << 1)) >> 1, -tc0), tc0));
		if (aq < beta) pix[step] = (uint8_t)(q1 + min(max((q2 + ((p0 + q0 + 1) >> 1) - (q1 << 1)) >> 1, -tc0), tc0));
	} else {
		bool small = iabs_(p0 - q0) < ((alpha >> 2) + 2);
		if (ap < beta && small) {
			int p3 = pix[-4 * step];
			pix[-step] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
			pix[-2 * step] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
			pix[-3 * step] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
		} else pix[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
		if (aq < beta && small) {
			int q3 = pix[3 * step];
			pix[0] = (uint8_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
			pix[step] = (uint8_t)((p0 + q0 + q1 + q2 + 2) >> 2);
			pix[2 * step] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
		} else pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
	}
}
__device__ __forceinline__ void filter_chroma(uint8_t *pix, int step, int bs, int alpha, int beta, int tc0) {
	int p0 = pix[-step], p1 = pix[-2 * step], q0 = pix[0], q1 = pix[step];
	if (!(iabs_(p0 - q0) < alpha && iabs_(p1 - p0) < beta && iabs_(q1 - q0) < beta)) return;
	if (bs < 4) {
		int tc = tc0 + 1;
		int d = min(max((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc), tc);
		pix[-step] = (uint8_t)clip255(p0 + d); pix[0] = (uint8_t)clip255(q0 - d);
	} else { pix[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2); pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2); }
}

/* One warp per macroblock ROW: the warp walks its row left to right, so the left neighbour's samples
 * stay in shared memory and only the row above is a cross-warp dependency, published as a per-row
 * progress counter (value = epoch * 2048 + macroblocks finished).  Row y may process macroblock x once
 * row y-1 has finished x+1 (its left-edge filter touches columns 13..15 of macroblock x above us). */
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) e264_deblock_rows_kernel(PicJob J) {   /* round-1 kernel, kept for A/B runs (E264B_DBK_OLD=1) */
	TraceScope trace_(J, 3);
	__shared__ DbSmem smem[WARPS_PER_BLOCK];
	const int lane = threadIdx.x & 31;
	DbSmem *ds = &smem[threadIdx.x >> 5];
	const int nmb = J.w_mbs * J.h_mbs, W = J.w_mbs;
	volatile unsigned *progress = J.flags + nmb;
	const unsigned base = J.epoch * 2048u;
	uint8_t *dst = J.frames + (size_t)J.dst_slot * J.frame_bytes;
	const int cpl = J.stride_c >> 1;
	for (;;) {
		unsigned t = 0;
		if (lane == 0) t = atomicAdd(J.tickets + 1, 1u);
		t = __shfl_sync(0xffffffffu, t, 0);
		if (t >= (unsigned)J.h_mbs) break;
		const int mby = (int)t;
		uint8_t *Yrow = dst + (size_t)(mby * 16) * J.stride_y;
		uint8_t *Crow = dst + J.plane_y + (size_t)(mby * 8) * J.stride_c;
		/* prefetch the first macroblock of the row: samples and records */
		uint4 nl = make_uint4(0, 0, 0, 0); uint2 nc = make_uint2(0, 0);
		if (lane < 16) nl = *(const uint4 *)(Yrow + (size_t)lane * J.stride_y);
		else { int j = lane - 16; nc = *(const uint2 *)(Crow + (j >> 3) * cpl + (size_t)(j & 7) * J.stride_c); }
		if (lane < 12) ds->rq[lane] = __ldg((const uint4 *)(J.recs + mby * W) + lane);
		else if (lane < 24 && mby > 0) ds->rt[lane - 12] = __ldg((const uint4 *)(J.recs + (mby - 1) * W) + lane - 12);
		uint4 tl = make_uint4(0, 0, 0, 0); uint2 tc = make_uint2(0, 0); bool have_top = false;
		__syncwarp();
		for (int mbx = 0; mbx < W; mbx++) {
			const int mb = mby * W + mbx;
			const E264MbRec *q = (const E264MbRec *)ds->rq, *pL = (const E264MbRec *)ds->rl, *pT = (const E264MbRec *)ds->rt;
			const int qflags = q->flags;
			uint8_t *Y = Yrow + mbx * 16, *C = Crow + mbx * 8;
			/* current macroblock into the tile; columns -4..-1 were left there by the previous iteration */
			if (lane < 16) *(uint4 *)&DY(0, lane) = nl;
			else { int j = lane - 16; *(uint2 *)&DC_(j >> 3, 0, j & 7) = nc; }
			if (have_top) {
				if (lane < 4) *(uint4 *)&DY(0, lane - 4) = tl;
				else if (lane < 8) { int j = lane - 4; *(uint2 *)&DC_(j >> 1, 0, (j & 1) - 2) = tc; }
			}
			const bool had_top = have_top;
			/* software pipeline: everything the NEXT macroblock needs is requested now */
			uint4 nrec = make_uint4(0, 0, 0, 0);
			have_top = false;
			if (mbx + 1 < W) {
				if (lane < 16) nl = *(const uint4 *)(Y + 16 + (size_t)lane * J.stride_y);
				else { int j = lane - 16; nc = *(const uint2 *)(C + 8 + (j >> 3) * cpl + (size_t)(j & 7) * J.stride_c); }
				if (lane < 12) nrec = __ldg((const uint4 *)(J.recs + mb + 1) + lane);
				else if (lane < 24 && mby > 0) nrec = __ldg((const uint4 *)(J.recs + mb + 1 - W) + lane - 12);
				if (mby > 0) {
					unsigned need = base + (unsigned)min(mbx + 3, W);
					int okp = 0;
					if (lane == 0) { okp = (int)(progress[mby - 1] - need) >= 0; if (okp) __threadfence(); }
					okp = __shfl_sync(0xffffffffu, okp, 0);
					if (okp) {
						have_top = true;
						if (lane < 4) tl = __ldcg((const uint4 *)(Y + 16 - (size_t)(4 - lane) * J.stride_y));
						else if (lane < 8) { int j = lane - 4, pl = j >> 1, r = (j & 1) - 2; tc = __ldcg((const uint2 *)(C + 8 + pl * cpl + (ptrdiff_t)r * J.stride_c)); }
					}
				}
			}
			if (qflags & MBF_DEBLOCK) {
				const E264SliceRec *sr = J.slices + q->slice_idx;
				const bool fl = qflags & MBF_EDGE_L, ft = qflags & MBF_EDGE_T, t8 = qflags & MBF_T8x8;
				{
					int dir = lane >> 4, e = (lane >> 2) & 3, k = lane & 3, bs = 0;
					const E264MbRec *p = q;
					bool on = true;
					if (e == 0) { on = dir ? ft : fl; p = dir ? pT : pL; }
					if (on) {
						int qx = dir ? k : e, qy = dir ? e : k;
						int px_ = dir ? k : (e ? e - 1 : 3), py_ = dir ? (e ? e - 1 : 3) : k;
						bs = bs_pair(p, blk_z(px_, py_), q, blk_z(qx, qy), e == 0);
					}
					ds->bs[lane] = (int8_t)bs;
				}
				if (lane < 9) {
					int pl = lane / 3, kind = lane % 3;
					const E264MbRec *p = kind == 0 ? q : kind == 1 ? pL : pT;
					if ((kind == 1 && !fl) || (kind == 2 && !ft)) p = q;
					int qpav = (p->qp[pl] + q->qp[pl] + 1) >> 1;
					int ia = min(max(qpav + sr->filter_offset_a, 0), 51), ib = min(max(qpav + sr->filter_offset_b, 0), 51);
					ds->alpha[pl][kind] = h264_alpha[ia]; ds->beta[pl][kind] = h264_beta[ib]; ds->ia[pl][kind] = (uint8_t)ia;
				}
				if (ft && !had_top) {
					/* the row above must have finished macroblock mbx+1 (or its whole row) */
					if (lane == 0) {
						unsigned need = base + (unsigned)min(mbx + 2, W), spins = 0;
						while ((int)(progress[mby - 1] - need) < 0) { __nanosleep(32); if ((++spins & 63) == 0 && (*(const volatile unsigned *)J.err || spins > (1u << 22))) { atomicExch(J.err, 1u); break; } }
						__threadfence();
					}
					__syncwarp();
					if (lane < 4) *(uint4 *)&DY(0, lane - 4) = __ldcg((const uint4 *)(Y - (size_t)(4 - lane) * J.stride_y));
					else if (lane < 8) { int j = lane - 4, pl = j >> 1, r = (j & 1) - 2; *(uint2 *)&DC_(pl, 0, r) = __ldcg((const uint2 *)(C + pl * cpl + (ptrdiff_t)r * J.stride_c)); }
				}
				__syncwarp();
				for (int dir = 0; dir < 2; dir++) {
					for (int e = 0; e < 4; e++) {
						int kind = e ? 0 : 1 + dir;
						if (lane < 16) {
							if (!(t8 && (e & 1))) {
								int b = ds->bs[dir * 16 + e * 4 + (lane >> 2)];
								if (b) {
									uint8_t *pix = dir ? &DY(lane, e * 4) : &DY(e * 4, lane);
									filter_luma(pix, dir ? 32 : 1, b, ds->alpha[0][kind], ds->beta[0][kind], b < 4 ? h264_tc0[ds->ia[0][kind]][b - 1] : 0);
								}
							}
						} else if (!(e & 1)) {
							int j = lane - 16, pl = j >> 3, k = j & 7;
							int b = ds->bs[dir * 16 + e * 4 + (k >> 1)];
							if (b) {
								uint8_t *pix = dir ? &DC_(pl, k, e * 2) : &DC_(pl, e * 2, k);
								filter_chroma(pix, dir ? 16 : 1, b, ds->alpha[1 + pl][kind], ds->beta[1 + pl][kind], b < 4 ? h264_tc0[ds->ia[1 + pl][kind]][b - 1] : 0);
							}
						}
						__syncwarp();
					}
				}
				/* write back: the macroblock, 3 columns of the left neighbour, 3 rows (1 for chroma) of the top neighbour */
				if (lane < 16) *(uint4 *)(Y + (size_t)lane * J.stride_y) = *(const uint4 *)&DY(0, lane);
				else { int j = lane - 16, pl = j >> 3, row = j & 7; *(uint2 *)(C + pl * cpl + (size_t)row * J.stride_c) = *(const uint2 *)&DC_(pl, 0, row); }
				if (fl) {
					if (lane < 16) { uint8_t *d = Y + (size_t)lane * J.stride_y; d[-3] = DY(-3, lane); d[-2] = DY(-2, lane); d[-1] = DY(-1, lane); }
					else { int j = lane - 16, pl = j >> 3, row = j & 7; C[pl * cpl + (size_t)row * J.stride_c - 1] = DC_(pl, -1, row); }
				}
				if (ft) {
					if (lane < 3) *(uint4 *)(Y - (size_t)(lane + 1) * J.stride_y) = *(const uint4 *)&DY(0, -1 - lane);
					else if (lane < 5) { int pl = lane - 3; *(uint2 *)(C + pl * cpl - J.stride_c) = *(const uint2 *)&DC_(pl, 0, -1); }
				}
			}
			__syncwarp();
			/* carry the last 4 (2) columns over as the next macroblock's left neighbour */
			if (lane < 16) *(uint32_t *)&DY(-4, lane) = *(const uint32_t *)&DY(12, lane);
			else { int j = lane - 16; *(uint16_t *)&DC_(j >> 3, -2, j & 7) = *(const uint16_t *)&DC_(j >> 3, 6, j & 7); }
			if (lane == 0) { __threadfence(); progress[mby] = base + (unsigned)mbx + 1u; }
			/* rotate the record buffers: current -> left, prefetched -> current / top */
			if (lane < 12) { ds->rl[lane] = ds->rq[lane]; ds->rq[lane] = nrec; }
			else if (lane < 24) ds->rt[lane - 12] = nrec;
			__syncwarp();
		}
	}
}
