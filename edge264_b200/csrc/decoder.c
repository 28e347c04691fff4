/* decoder.c — the edge264 C API over the record-producing parser and a reconstruction backend.
 *
 * Host-side counterpart of the reference's edge264.c + edge264_headers.c (reference: API
 * edge264.c:87-415; SPS/PPS/slice header/POC/ref lists/MMCO/bumping edge264_headers.c:61-141,
 * 611-1290, 1343-2059).  Everything is re-implemented from ITU-T H.264 with the reference's
 * observable behaviour (return codes, output order, frame layout, FrameId numbering) as the
 * contract; no pixel is touched here: pictures are handed to the backend as records.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "dec.h"
#include "h264_tables.h"

int e264_parse_slice_data(SliceCtx *s);   /* slice_dec.c */

/* optional host-side time accounting (E264_HOST_PROFILE=1): where decode_NAL/get_frame spend their time */
#include <time.h>
static int prof_on = -1;
static double prof_t[6]; static long prof_n[6];
static const char *const prof_names[6] = {"parse_slice_data", "acquire_staging", "submit", "wait", "clear_mbinfo", "headers+dpb"};
static inline double prof_now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
#define PROF_BEGIN double prof_t0_ = prof_on > 0 ? prof_now() : 0
#define PROF_END(i) do { if (prof_on > 0) { prof_t[i] += prof_now() - prof_t0_; prof_n[i]++; } } while (0)

/* ------------------------------------------------------------------------------------------ */
/* start codes (reference edge264.c:87-119: returns a pointer to the 00 00 01 / 00 00 00 01)     */
/* ------------------------------------------------------------------------------------------ */
const uint8_t *edge264_find_start_code(const uint8_t *buf, const uint8_t *end, int four_byte) {
	four_byte = four_byte != 0;
	const uint8_t *p = buf + four_byte;   /* the reference searches a 001 from buf+four_byte and backs up */
	if (p >= end) return end;
	for (;;) {
		/* find next 0x01 at or after p+2 (position of the '1' of a 001 that starts at >= p) */
		const uint8_t *s = p + 2;
		if (s >= end) return end;
		const uint8_t *one = (const uint8_t *)memchr(s, 1, (size_t)(end - s));
		if (!one) return end;
		if (one[-1] == 0 && one[-2] == 0) {
			const uint8_t *res = one - 2 - four_byte;
			if (*res == 0) return res < end ? res : end;
		}
		p = one - 1;   /* continue after this 0x01: next candidate '1' position is > one */
		if (p + 2 <= one) p = one - 1;
	}
}

/* ------------------------------------------------------------------------------------------ */
/* parameter sets                                                                               */
/* ------------------------------------------------------------------------------------------ */
static const uint32_t max_dpb_mbs_by_level[64] = {   /* Table A-1 MaxDpbMbs, indexed by level_idc */
	396, 396, 396, 396, 396, 396, 396, 396, 396, 396, 396, 900, 2376, 2376, 2376, 2376, 2376, 2376, 2376, 2376, 2376,
	4752, 8100, 8100, 8100, 8100, 8100, 8100, 8100, 8100, 8100, 18000, 20480, 32768, 32768, 32768, 32768, 32768, 32768,
	32768, 32768, 32768, 34816, 110400, 110400, 110400, 110400, 110400, 110400, 110400, 110400, 184320, 184320,
	696320, 696320, 696320, 696320, 696320, 696320, 696320, 696320, 696320, 696320, 0xffffffffu};

/* scaling_list(): returns 0 = explicit list written to dst (raster), 1 = useDefaultScalingMatrixFlag */
static int parse_scaling_list(BitReader *b, uint8_t *dst, int n, const uint8_t *scan) {
	int last = 8, next = 8;
	for (int j = 0; j < n; j++) {
		if (next != 0) {
			int delta = br_se(b);
			next = (last + delta + 256) & 255;
			if (j == 0 && next == 0) return 1;
		}
		dst[scan[j]] = (uint8_t)(next == 0 ? last : next);
		last = dst[scan[j]];
	}
	return 0;
}

/* The reference keeps not-transmitted PPS lists as zeros and substitutes the SPS lists at slice time
 * (reference headers.c:903-915, 1343-1413); we reproduce exactly that observable behaviour:
 *  lists: 6 of 4x4 then 2 of 8x8; `w4`/`w8` hold the fall-back values on entry. */
static void parse_scaling_lists(BitReader *b, uint8_t w4[6][16], uint8_t w8[2][64], int with8x8) {
	uint8_t fb[16]; memcpy(fb, w4[0], 16);
	const uint8_t *def = h264_default4x4_intra;
	for (int i = 0; i < 6; i++) {
		if (i == 3) { memcpy(fb, w4[3], 16); def = h264_default4x4_inter; }
		if (!br_u1(b)) memcpy(w4[i], fb, 16);
		else {
			if (parse_scaling_list(b, w4[i], 16, h264_zigzag4x4)) memcpy(w4[i], def, 16);
			memcpy(fb, w4[i], 16);
		}
	}
	if (!with8x8) return;
	for (int i = 0; i < 2; i++) {
		if (!br_u1(b)) continue;
		if (parse_scaling_list(b, w8[i], 64, h264_zigzag8x8)) memcpy(w8[i], i ? h264_default8x8_inter : h264_default8x8_intra, 64);
	}
}

static void skip_hrd(BitReader *b) {
	int cpb_cnt = br_ue_i(b, 32) + 1;
	br_u(b, 8);
	for (int i = 0; i < cpb_cnt && i < 32; i++) { br_ue(b); br_ue(b); br_u1(b); }
	br_u(b, 20);
}

static int parse_sps(Edge264Decoder *d, BitReader *b, SPS *out) {
	SPS s; memset(&s, 0, sizeof(s));
	int ret = 0;
	memset(s.sl4x4, 16, sizeof(s.sl4x4)); memset(s.sl8x8, 16, sizeof(s.sl8x8));
	s.profile_idc = br_u(b, 8); s.constraint_flags = br_u(b, 8); s.level_idc = br_u(b, 8);
	br_ue(b);   /* seq_parameter_set_id: ignored like the reference (headers.c:1851) */
	s.log2_max_poc_lsb = 16;
	if (s.profile_idc != 66 && s.profile_idc != 77 && s.profile_idc != 88) {
		int chroma_format_idc = br_ue_i(b, 4);
		if (chroma_format_idc != 1) { ret = ENOTSUP; if (chroma_format_idc == 3) br_u1(b); }
		if (br_ue(b) != 0) ret = ENOTSUP;   /* bit_depth_luma_minus8 */
		if (br_ue(b) != 0) ret = ENOTSUP;
		if (br_u1(b)) ret = ENOTSUP;        /* qpprime_y_zero_transform_bypass_flag */
		if (br_u1(b)) {                     /* seq_scaling_matrix_present_flag */
			s.scaling_present = 1;
			memcpy(s.sl4x4[0], h264_default4x4_intra, 16); memcpy(s.sl4x4[3], h264_default4x4_inter, 16);
			memcpy(s.sl8x8[0], h264_default8x8_intra, 64); memcpy(s.sl8x8[1], h264_default8x8_inter, 64);
			if (chroma_format_idc == 3) return ENOTSUP;
			parse_scaling_lists(b, s.sl4x4, s.sl8x8, 1);
		}
	}
	s.log2_max_frame_num = br_ue_i(b, 13) + 4;
	s.poc_type = br_ue_i(b, 3);
	if (s.log2_max_frame_num > 16 || s.poc_type > 2) return EBADMSG;
	if (s.poc_type == 0) { s.log2_max_poc_lsb = br_ue_i(b, 13) + 4; if (s.log2_max_poc_lsb > 16) return EBADMSG; }
	else if (s.poc_type == 1) {
		s.delta_pic_order_always_zero_flag = br_u1(b);
		s.offset_for_non_ref_pic = br_se(b);
		s.offset_for_top_to_bottom_field = br_se(b);
		s.num_ref_frames_in_poc_cycle = br_ue_i(b, 256);
		if (s.num_ref_frames_in_poc_cycle > 255) return EBADMSG;
		for (int i = 0; i < s.num_ref_frames_in_poc_cycle; i++) s.offset_for_ref_frame[i] = br_se(b);
	}
	int max_num_ref_frames = br_ue_i(b, 17);
	s.gaps_allowed = br_u1(b);
	s.width_mbs = br_ue_i(b, 4096) + 1;
	s.height_mbs = br_ue_i(b, 4096) + 1;
	if (s.width_mbs > 1023 || s.height_mbs > 1055 || max_num_ref_frames > 16) return EBADMSG;
	int frame_mbs_only = br_u1(b);
	if (!frame_mbs_only) { ret = ENOTSUP; br_u1(b); }
	s.direct_8x8_inference = br_u1(b);
	unsigned lvl = s.level_idc < 63 ? s.level_idc : 63;
	int max_dpb_frames = (int)(max_dpb_mbs_by_level[lvl] / (unsigned)(s.width_mbs * s.height_mbs));
	if (max_dpb_frames > 16) max_dpb_frames = 16;
	s.max_num_ref_frames = max_num_ref_frames < max_dpb_frames ? max_num_ref_frames : max_dpb_frames;
	int intra_profile = (s.profile_idc == 44 || s.profile_idc == 86 || s.profile_idc == 100 || s.profile_idc == 110 ||
	                     s.profile_idc == 122 || s.profile_idc == 244) && (s.constraint_flags & (1 << 4));
	if (intra_profile) { s.max_num_reorder_frames = 0; s.max_dec_frame_buffering = s.max_num_ref_frames; }
	else s.max_num_reorder_frames = s.max_dec_frame_buffering = max_dpb_frames;
	if (br_u1(b)) {   /* frame_cropping_flag; 4:2:0 frame: units of 2 luma samples */
		/* out-of-range offsets are clamped, not rejected, like the reference's bounded get_ue16 (headers.c:1975-1983) */
		int limx = s.width_mbs * 8 - 1, limy = s.height_mbs * 8 - 1;
		int v = br_ue_i(b, 1 << 20); s.crop[0] = (v < 0 || v > limx ? limx : v) * 2;
		v = br_ue_i(b, 1 << 20); s.crop[1] = (v < 0 || v > limx - s.crop[0] / 2 ? limx - s.crop[0] / 2 : v) * 2;
		v = br_ue_i(b, 1 << 20); s.crop[2] = (v < 0 || v > limy ? limy : v) * 2;
		v = br_ue_i(b, 1 << 20); s.crop[3] = (v < 0 || v > limy - s.crop[2] / 2 ? limy - s.crop[2] / 2 : v) * 2;
	}
	if (br_u1(b)) {   /* vui_parameters (E.1.1): walked only to reach bitstream_restriction */
		if (br_u1(b)) { if (br_u(b, 8) == 255) { br_u(b, 16); br_u(b, 16); } }
		if (br_u1(b)) br_u1(b);
		if (br_u1(b)) { br_u(b, 4); if (br_u1(b)) br_u(b, 24); }
		if (br_u1(b)) { br_ue(b); br_ue(b); }
		if (br_u1(b)) { br_u(b, 32); br_u(b, 32); br_u1(b); }
		int nal_hrd = br_u1(b); if (nal_hrd) skip_hrd(b);
		int vcl_hrd = br_u1(b); if (vcl_hrd) skip_hrd(b);
		if (nal_hrd || vcl_hrd) br_u1(b);
		br_u1(b);   /* pic_struct_present_flag */
		if (br_u1(b)) {
			br_u1(b); br_ue(b); br_ue(b); br_ue(b); br_ue(b);
			int reorder = br_ue_i(b, 17), buffering = br_ue_i(b, 17);
			if (reorder > 16 || buffering > 16) return EBADMSG;
			s.max_dec_frame_buffering = buffering > s.max_num_ref_frames ? buffering : s.max_num_ref_frames;
			s.max_num_reorder_frames = reorder < s.max_dec_frame_buffering ? reorder : s.max_dec_frame_buffering;
		}
	}
	if (b->overrun) return EBADMSG;
	(void)d;
	s.valid = 1;
	*out = s;
	return ret;
}

static int parse_pps(Edge264Decoder *d, BitReader *b) {
	PPS p; memset(&p, 0, sizeof(p));
	int ret = 0;
	unsigned id = br_ue(b);
	if (id >= E264_MAX_PPS || id >= 4) ret = ENOTSUP;   /* the reference supports 4 PPS ids (headers.c:1435) */
	br_ue(b);
	p.entropy_coding_mode = br_u1(b);
	p.bottom_field_pic_order_present = br_u1(b);
	if (br_ue(b) != 0) return ENOTSUP;   /* slice groups: cannot parse further */
	p.num_ref_idx_default[0] = br_ue_i(b, 32) + 1; p.num_ref_idx_default[1] = br_ue_i(b, 32) + 1;
	if (p.num_ref_idx_default[0] > 32 || p.num_ref_idx_default[1] > 32) return EBADMSG;
	p.weighted_pred_flag = br_u1(b); p.weighted_bipred_idc = br_u(b, 2);
	int q = br_se(b); if (q < -26 || q > 25) return EBADMSG;
	p.pic_init_qp = 26 + q;
	br_se(b);
	p.chroma_qp_index_offset[0] = p.chroma_qp_index_offset[1] = br_se(b);
	if (p.chroma_qp_index_offset[0] < -12 || p.chroma_qp_index_offset[0] > 12) return EBADMSG;
	p.deblocking_filter_control_present = br_u1(b);
	if (br_u1(b)) ret = ENOTSUP;   /* constrained_intra_pred_flag */
	if (br_u1(b)) ret = ENOTSUP;   /* redundant_pic_cnt_present_flag */
	int has_matrix = 0;
	if (br_more_rbsp_data(b)) {
		p.transform_8x8_mode = br_u1(b);
		has_matrix = br_u1(b);
		if (has_matrix) parse_scaling_lists(b, p.sl4x4, p.sl8x8, p.transform_8x8_mode);   /* zeros = "take the SPS list" */
		p.chroma_qp_index_offset[1] = br_se(b);
		if (p.chroma_qp_index_offset[1] < -12 || p.chroma_qp_index_offset[1] > 12) return EBADMSG;
	}
	if (b->overrun || p.weighted_bipred_idc == 3) return EBADMSG;
	p.valid = 1 + has_matrix;
	if (ret == 0) d->pps[id] = p;
	return ret;
}

/* scaling lists in force for a slice (reference merge rule, headers.c:903-915) */
static void merge_scaling(const SPS *s, const PPS *p, E264SliceRec *sr) {
	if (p->valid == 2) {
		for (int i = 0; i < 6; i++) {
			const uint8_t *fb = s->sl4x4[i < 3 ? 0 : 3];
			for (int k = 0; k < 16; k++) sr->scaling4x4[i][k] = p->sl4x4[i][k] ? p->sl4x4[i][k] : fb[k];
		}
		for (int i = 0; i < 2; i++) for (int k = 0; k < 64; k++) sr->scaling8x8[i][k] = p->sl8x8[i][k] ? p->sl8x8[i][k] : s->sl8x8[i][k];
	} else {
		memcpy(sr->scaling4x4, s->sl4x4, sizeof(sr->scaling4x4));
		memcpy(sr->scaling8x8, s->sl8x8, sizeof(sr->scaling8x8));
	}
}

/* ------------------------------------------------------------------------------------------ */
/* host mirrors and output queue                                                                */
/* ------------------------------------------------------------------------------------------ */
static int hostbuf_acquire(Edge264Decoder *d) {
	for (int i = 0; i < E264_MAX_HOSTBUFS; i++) if (d->hb[i].state == 0 && d->hb[i].p) return i;
	for (int i = 0; i < E264_MAX_HOSTBUFS; i++) if (d->hb[i].state == 0) {
		size_t bytes = (size_t)d->frame_bytes + 64;
		if (d->alloc_cb) {
			void *smp = NULL, *mbs = NULL;
			/* reference sizes (headers.c:113-117): samples = plane_Y + plane_C + 16, mbs = 304 * ((W+1)*H - 1) */
			d->alloc_cb(&smp, (unsigned)(d->plane_y + d->plane_c + 16), &mbs, (unsigned)(304 * ((d->w_mbs + 1) * d->h_mbs - 1)), ENOMEM, d->alloc_arg);
			if (!smp) return -1;
			d->hb[i].p = (uint8_t *)smp; d->hb[i].mbs = mbs;
		} else {
			if (d->n_threads) pthread_mutex_lock(&d->be_lock);
			d->hb[i].p = (uint8_t *)d->be->host_alloc(d->be_ctx, bytes);
			if (d->n_threads) pthread_mutex_unlock(&d->be_lock);
			if (!d->hb[i].p) return -1;
		}
		return i;
	}
	return -1;
}
static void hostbufs_free_all(Edge264Decoder *d) {
	for (int i = 0; i < E264_MAX_HOSTBUFS; i++) {
		if (d->hb[i].p) {
			if (d->alloc_cb) d->free_cb(d->hb[i].p, d->hb[i].mbs, d->alloc_arg);
			else d->be->host_free(d->be_ctx, d->hb[i].p);
		}
		memset(&d->hb[i], 0, sizeof(d->hb[i]));
	}
	d->outq_n = 0; d->pending_release = -1;
}

/* DPB bookkeeping mirrors the reference's bit sets (edge264_internal.h:1217-1222):
 *   ref            <-> prev_short_term_frames / prev_long_term_frames
 *   needed_for_output <-> to_get_frames & ~output_frames */
static int dpb_fullness(Edge264Decoder *d) {
	int n = 0;
	for (int i = 0; i < d->n_slots; i++) if (d->pics[i].in_use && (d->pics[i].ref || d->pics[i].needed_for_output)) n++;
	return n;
}
static int waiting_for_output(Edge264Decoder *d) {
	int n = 0;
	for (int i = 0; i < d->n_slots; i++) if (d->pics[i].in_use && d->pics[i].needed_for_output) n++;
	return n;
}
static void slot_release_if_unused(Edge264Decoder *d, int i) {
	Pic *p = &d->pics[i];
	if (p->in_use && !p->ref && !p->needed_for_output && i != d->cur) p->in_use = 0;
}
/* bumping: output the waiting picture with the lowest POC (reference bump_frame, headers.c:78-95) */
static int bump_frame(Edge264Decoder *d, int ignore_slot) {
	int best = -1;
	for (int i = 0; i < d->n_slots; i++) {
		Pic *p = &d->pics[i];
		if (!p->in_use || !p->needed_for_output || i == ignore_slot) continue;
		if (best < 0 || p->poc_top < d->pics[best].poc_top) best = i;
	}
	if (best < 0) return 0;
	Pic *p = &d->pics[best];
	p->needed_for_output = 0;
	if (p->host_buf >= 0) {
		d->hb[p->host_buf].state = 2; d->outq[d->outq_n++] = p->host_buf;
		if (d->hb[p->host_buf].submitted) p->host_buf = -1;
	}
	slot_release_if_unused(d, best);
	return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* picture completion                                                                           */
/* ------------------------------------------------------------------------------------------ */
static void apply_marking(Edge264Decoder *d);

/* Macroblocks that no slice delivered (lost or damaged slices): give them a neutral, self-consistent record so
 * that every picture the backend sees is complete — copy of the co-located samples of the most recent reference
 * picture (a P_Skip with zero motion), or DC intra prediction when there is none.  The reference blends an
 * error-probability-weighted intra DC / re-runs P_Skip (recover_slice, edge264_headers.c:295-430); matching its
 * concealed samples is not attempted, only a deterministic, safe picture. */
static void conceal_missing(Edge264Decoder *d, PicBuild *pb) {
	const int total = d->w_mbs * d->h_mbs;
	const int ref = pb->conceal_ref;
	if (pb->n_slices == 0) {   /* not one slice arrived: a neutral slice record for the concealed macroblocks to point to */
		E264SliceRec *sr = &pb->slices[0];
		memset(sr, 0, sizeof(*sr));
		memset(sr->scaling4x4, 16, sizeof(sr->scaling4x4)); memset(sr->scaling8x8, 16, sizeof(sr->scaling8x8));
		pb->n_slices = 1;
	}
	for (int a = 0; a < total; a++) {
		if (pb->mbi[a].slice_id) continue;
		E264MbRec *r = pb->recs + a;
		memset(r, 0, sizeof(*r));
		r->qp[0] = r->qp[1] = r->qp[2] = 26;
		memset(r->ref_idx, -1, sizeof(r->ref_idx)); memset(r->ref_pic, -1, sizeof(r->ref_pic));
		if (ref >= 0) {
			r->kind = MBK_INTER; r->flags = MBF_SKIP;
			for (int i8 = 0; i8 < 4; i8++) { r->ref_idx[0][i8] = 0; r->ref_pic[0][i8] = (int8_t)ref; }
		} else {
			r->kind = MBK_I16x16; r->i16_mode = IMODE(2, 1 | 2 | 8); r->chroma_mode = IMODE(0, 1 | 2 | 8);
		}
	}
}

/* the frame slot a lost macroblock is copied from: the most recent reference picture (header-time state) */
static int conceal_reference(Edge264Decoder *d) {
	int ref = -1, best = -1;
	for (int i = 0; i < d->n_slots; i++) if (i != d->cur && d->pics[i].in_use && d->pics[i].ref && !d->pics[i].nonexisting && d->pics[i].uid > best) { best = d->pics[i].uid; ref = i; }
	return ref;
}

/* data side of a finished picture: conceal what no slice delivered and list the intra macroblocks (build_finalize: the
 * records are complete afterwards), then hand the picture to the backend (build_submit).  Both run in the calling
 * thread (synchronous mode) or in a worker once every slice is parsed. */
static void build_finalize(Edge264Decoder *d, PicBuild *pb, E264PicDesc *pd) {
	if (pb->mbs_done < d->w_mbs * d->h_mbs) conceal_missing(d, pb);
	memset(pd, 0, sizeof(*pd));
	pd->width_mbs = d->w_mbs; pd->height_mbs = d->h_mbs; pd->stride_y = d->stride_y; pd->stride_c = d->stride_c;
	pd->plane_y = d->plane_y; pd->frame_bytes = d->frame_bytes; pd->dst_slot = pb->slot;
	pd->n_slices = pb->n_slices; pd->n_coefs = (int32_t)pb->n_coefs; pd->any_deblock = pb->any_deblock; pd->staging = pb->staging;
	/* intra macroblocks in raster order: the device draws them from this list, so a waiting macroblock only ever waits for earlier entries */
	const int total = d->w_mbs * d->h_mbs; int n = 0;
	for (int a = 0; a < total; a++) if (pb->recs[a].kind != MBK_INTER) pb->intra_list[n++] = (uint32_t)a;
	pd->n_intra = n;
}
static int build_submit(Edge264Decoder *d, PicBuild *pb, const E264PicDesc *pd, uint64_t *ticket) {
	*ticket = 0;
	if (pb->host_buf < 0) return 0;
	int r;
	{ PROF_BEGIN; r = d->be->submit(d->be_ctx, pd, d->hb[pb->host_buf].p, ticket); PROF_END(2); }
	return r ? EIO : 0;
}

static void close_picture_threaded(Edge264Decoder *d);

static int finish_picture(Edge264Decoder *d) {
	if (d->cur < 0) return 0;
	Pic *p = &d->pics[d->cur];
	int ret = 0;
	if (d->n_threads) close_picture_threaded(d);
	else {
		PicBuild *pb = d->pb;
		pb->conceal_ref = conceal_reference(d);
		if (p->host_buf >= 0) {
			uint64_t ticket = 0; E264PicDesc pd;
			build_finalize(d, pb, &pd);
			ret = build_submit(d, pb, &pd, &ticket);
			d->hb[p->host_buf].ticket = ticket; __atomic_store_n(&d->hb[p->host_buf].submitted, 1, __ATOMIC_RELEASE);
		}
		pb->in_use = 0;
	}
	d->pb = NULL;
	if (p->host_buf >= 0 && !p->needed_for_output) p->host_buf = -1;   /* already in the output queue */
	apply_marking(d);
	int c = d->cur;
	d->cur = -1;
	slot_release_if_unused(d, c);
	return ret;
}

/* ------------------------------------------------------------------------------------------ */
/* worker threads: several pictures parsed at the same time                                      */
/* ------------------------------------------------------------------------------------------ */
/* The calling thread parses headers, runs the decoded-picture-buffer logic and queues slices; workers parse slice
 * data.  One worker owns a picture at a time (its slices share the coefficient pool and run in order); pictures are
 * handed to the backend strictly in decoding order.  A B slice waits until its co-located picture is completely
 * parsed (its records are the motion data direct prediction reads); nothing else depends on another picture at parse
 * time — pixels are the device's business.  Reference counterpart: task dependencies + worker_loop,
 * edge264_internal.h:1196-1226, edge264_headers.c:450-603. */
static PicBuild *build_acquire(Edge264Decoder *d, int slot) {
	if (!d->n_threads) { PicBuild *pb = &d->builds[0]; pb->in_use = 1; return pb; }
	pthread_mutex_lock(&d->lock);
	PicBuild *pb = NULL;
	for (;;) {
		/* the slot's records may still be written by its previous picture's parser or read as co-located motion */
		if (d->slot_users[slot] == 0) for (int i = 0; i < d->n_builds && !pb; i++) if (!d->builds[i].in_use) pb = &d->builds[i];
		if (pb) break;
		pthread_cond_wait(&d->done_cv, &d->lock);
	}
	MbInfo *mbi = pb->mbi;
	memset(pb, 0, sizeof(*pb));
	pb->mbi = mbi; pb->in_use = 1; pb->seq = d->next_seq++;
	d->slot_users[slot]++; d->slot_build[slot] = pb; d->slot_build_seq[slot] = pb->seq;
	pthread_mutex_unlock(&d->lock);
	return pb;
}
static void slice_enqueue(Edge264Decoder *d, PicBuild *pb, SliceJob *job, int col_slot) {
	pthread_mutex_lock(&d->lock);
	job->next = NULL; job->dep = NULL; job->dep_seq = 0; job->col_slot = col_slot;
	if (col_slot >= 0) {
		d->slot_users[col_slot]++;
		if (d->slot_build[col_slot] && d->slot_build[col_slot] != pb) { job->dep = d->slot_build[col_slot]; job->dep_seq = d->slot_build_seq[col_slot]; }
	}
	if (pb->tail) pb->tail->next = job; else pb->head = job;
	pb->tail = job;
	pthread_cond_broadcast(&d->work_cv);
	pthread_mutex_unlock(&d->lock);
}
static void close_picture_threaded(Edge264Decoder *d) {
	pthread_mutex_lock(&d->lock);
	d->pb->conceal_ref = conceal_reference(d);
	d->pb->closed = 1;
	pthread_cond_broadcast(&d->work_cv);
	pthread_mutex_unlock(&d->lock);
}
/* wait until every picture has been parsed and handed to the backend (flush, end of stream, format change) */
static void threads_drain(Edge264Decoder *d) {
	if (!d->n_threads) return;
	pthread_mutex_lock(&d->lock);
	for (;;) {
		int busy = 0;
		for (int i = 0; i < d->n_builds; i++) busy |= d->builds[i].in_use;
		if (!busy) break;
		pthread_cond_wait(&d->done_cv, &d->lock);
	}
	pthread_mutex_unlock(&d->lock);
}
static inline int job_ready(const SliceJob *j) { return !j->dep || j->dep->seq != j->dep_seq || j->dep->parsed || !j->dep->in_use; }
static void *worker_main(void *arg) {
	Edge264Decoder *d = (Edge264Decoder *)arg;
	pthread_mutex_lock(&d->lock);
	for (;;) {
		PicBuild *pb = NULL;
		while (!d->stop) {
			for (int i = 0; i < d->n_builds; i++) {     /* the oldest picture with something to do */
				PicBuild *b = &d->builds[i];
				if (!b->in_use || b->running || b->parsed) continue;
				if (!(b->head ? job_ready(b->head) : b->closed)) continue;
				if (!pb || b->seq < pb->seq) pb = b;
			}
			if (pb) break;
			pthread_cond_wait(&d->work_cv, &d->lock);
		}
		if (d->stop) break;
		pb->running = 1;
		while (pb->head && job_ready(pb->head)) {
			SliceJob *job = pb->head;
			pb->head = job->next; if (!pb->head) pb->tail = NULL;
			pthread_mutex_unlock(&d->lock);
			job->sc.n_coefs = pb->n_coefs;      /* the slices of a picture share one coefficient pool */
			int n = e264_parse_slice_data(&job->sc);
			pb->n_coefs = job->sc.n_coefs; pb->n_intra += job->sc.n_intra;
			if (n > 0) pb->mbs_done += n;
			if (job->sc.error) pb->error = job->sc.error;
			free(job->rbsp);
			pthread_mutex_lock(&d->lock);
			if (job->col_slot >= 0) { d->slot_users[job->col_slot]--; pthread_cond_broadcast(&d->done_cv); }
			free(job);
		}
		if (!pb->head && pb->closed) {
			pthread_mutex_unlock(&d->lock);
			build_finalize(d, pb, &pb->pd);
			pthread_mutex_lock(&d->lock);
			pb->parsed = 1; pb->running = 0;      /* its records are complete: B pictures may read them */
			pthread_cond_broadcast(&d->work_cv);
			/* hand parsed pictures to the backend in decoding order; whoever completes the oldest one also sends the
			 * ones that were waiting behind it, so no worker ever sleeps on the order */
			while (!d->submitting) {
				PicBuild *q = NULL;
				for (int i = 0; i < d->n_builds; i++) if (d->builds[i].in_use && d->builds[i].parsed && d->builds[i].seq == d->submit_seq) q = &d->builds[i];
				if (!q) break;
				uint64_t ticket = 0;
				d->submitting = 1;
				pthread_mutex_unlock(&d->lock);
				pthread_mutex_lock(&d->be_lock);
				build_submit(d, q, &q->pd, &ticket);
				pthread_mutex_unlock(&d->be_lock);
				pthread_mutex_lock(&d->lock);
				if (q->host_buf >= 0) { d->hb[q->host_buf].ticket = ticket; __atomic_store_n(&d->hb[q->host_buf].submitted, 1, __ATOMIC_RELEASE); }
				d->submit_seq++; d->submitting = 0;
				d->slot_users[q->slot]--;
				q->in_use = 0;
				pthread_cond_broadcast(&d->done_cv);
			}
		} else pb->running = 0;               /* more slices to come, or the next one waits for its co-located picture */
	}
	pthread_mutex_unlock(&d->lock);
	return NULL;
}
static void threads_stop(Edge264Decoder *d) {
	if (!d->n_threads) return;
	threads_drain(d);
	pthread_mutex_lock(&d->lock);
	d->stop = 1;
	pthread_cond_broadcast(&d->work_cv); pthread_cond_broadcast(&d->done_cv);
	pthread_mutex_unlock(&d->lock);
	for (int i = 0; i < d->n_threads; i++) pthread_join(d->threads[i], NULL);
	d->n_threads = 0;
}

/* 8.2.5 decoded reference picture marking, applied when the picture is complete.  Mirrors the
 * reference's order of operations (headers.c:611-701): MMCO list, then sliding window whenever the
 * number of references reaches max_num_ref_frames, then the current picture is inserted. */
static void apply_marking(Edge264Decoder *d) {
	Pic *cur = &d->pics[d->cur];
	SliceHeader *h = &d->first_sh;
	if (!d->cur_nal_ref_idc) { cur->ref = 0; return; }
	int cur_long = 0;
	if (d->cur_idr) {
		for (int i = 0; i < d->n_slots; i++) if (i != d->cur && d->pics[i].in_use) { d->pics[i].ref = 0; slot_release_if_unused(d, i); }
		cur_long = h->long_term_reference_flag;
		cur->long_term_idx = 0;
	} else {
		for (int k = 0; k < h->n_mmco; k++) {
			int op = h->mmco[k].op;
			int target = -1;
			if (op == 1 || op == 3) {
				int fn = cur->frame_num - 1 - (int)h->mmco[k].a;
				for (int i = 0; i < d->n_slots; i++) if (i != d->cur && d->pics[i].in_use && d->pics[i].ref == 1 && d->pics[i].frame_num == fn) { target = i; d->pics[i].ref = 0; }
			}
			if (op == 2 || op == 3 || op == 4 || op == 6) {
				int idx = (int)(op == 3 ? h->mmco[k].b : h->mmco[k].a);
				int up = op == 4 ? INT_MAX : idx;
				if (op == 4) idx = (int)h->mmco[k].a;   /* max_long_term_frame_idx_plus1: drop idx >= a */
				for (int i = 0; i < d->n_slots; i++) if (i != d->cur && d->pics[i].in_use && d->pics[i].ref == 2 && d->pics[i].long_term_idx >= idx && d->pics[i].long_term_idx <= up) d->pics[i].ref = 0;
				if (op == 3 && target >= 0) { d->pics[target].ref = 2; d->pics[target].long_term_idx = idx; }
				if (op == 6) { cur_long = 1; cur->long_term_idx = idx; }
			}
			if (op == 5) {
				for (int i = 0; i < d->n_slots; i++) if (i != d->cur && d->pics[i].in_use) d->pics[i].ref = 0;
			}
		}
		for (int i = 0; i < d->n_slots; i++) slot_release_if_unused(d, i);
	}
	int nref = 0;
	for (int i = 0; i < d->n_slots; i++) if (i != d->cur && d->pics[i].in_use && d->pics[i].ref) nref++;
	if (nref >= d->sps.max_num_ref_frames) {
		int best = -1;
		for (int i = 0; i < d->n_slots; i++) if (i != d->cur && d->pics[i].in_use && d->pics[i].ref == 1 && (best < 0 || d->pics[i].frame_num < d->pics[best].frame_num)) best = i;
		if (best >= 0) { d->pics[best].ref = 0; slot_release_if_unused(d, best); }
	}
	cur->ref = cur_long ? 2 : 1;
	for (int k = 0; k < h->n_mmco; k++) if (h->mmco[k].op == 5) cur->frame_num = 0;   /* 8.2.1: the picture is now "frame_num 0" for everything that follows */
	d->prev_ref_frame_num = cur->frame_num;
	d->q_prev_ref_frame_num = d->q_cur_frame_num;
	for (int k = 0; k < h->n_mmco; k++) if (h->mmco[k].op == 5) d->q_prev_ref_frame_num = 0;
}

/* ------------------------------------------------------------------------------------------ */
/* sequence (re)configuration                                                                   */
/* ------------------------------------------------------------------------------------------ */
static int bump_all(Edge264Decoder *d) {
	if (d->cur >= 0) finish_picture(d);
	threads_drain(d);
	while (bump_frame(d, -1));
	if (d->outq_n) return ENOBUFS;
	/* frames the application still holds (borrowed, or handed out until the next decode_NAL) keep their buffers: a
	 * change of format must wait for them like the reference's `to_get_frames | output_frames` test (headers.c:2005-2007) */
	for (int i = 0; i < E264_MAX_HOSTBUFS; i++) if (d->hb[i].state == 3 && d->hb[i].borrowed) return ENOBUFS;
	return 0;
}

/* keep_numbering: the frame format is unchanged and only the frame pool grows — the reference does not clear its decoder
 * then (it compares size, crop and bit depth only, headers.c:2016-2024), so FrameId numbering and the parameter sets go on */
static int configure_sequence(Edge264Decoder *d, const SPS *s, int keep_numbering) {
	int w = s->width_mbs * 16, h = s->height_mbs * 16;
	d->w_mbs = s->width_mbs; d->h_mbs = s->height_mbs;
	d->stride_y = w; if (!(d->stride_y & 2047)) d->stride_y += 16;       /* reference headers.c:2027-2029 */
	d->stride_c = w; if (!(d->stride_c & 4095)) d->stride_c += 16;       /* headers.c:2035-2037 pads by 8; 16 keeps both chroma planes 8-byte aligned for vector stores and TMA */
	d->plane_y = d->stride_y * h; d->plane_c = d->stride_c * (h >> 1);
	d->frame_bytes = d->plane_y + d->plane_c + 16;
	if (d->n_threads) pthread_mutex_lock(&d->lock);      /* idle workers scan builds[0..n_builds) */
	d->n_builds = d->n_threads ? d->n_threads + 2 : 1;
	if (d->n_builds > E264_MAX_BUILDS) d->n_builds = E264_MAX_BUILDS;
	if (d->n_threads) pthread_mutex_unlock(&d->lock);
	/* a picture keeps its slot (and the slot's record buffer) until it is handed to the device: parsing ahead needs one spare slot per picture in flight */
	d->n_slots = s->max_num_ref_frames + 2 + (d->n_threads ? d->n_builds : 0);
	if (d->n_slots > E264_MAX_SLOTS) d->n_slots = E264_MAX_SLOTS;
	E264PicDesc g; memset(&g, 0, sizeof(g));
	g.width_mbs = d->w_mbs; g.height_mbs = d->h_mbs; g.stride_y = d->stride_y; g.stride_c = d->stride_c; g.plane_y = d->plane_y; g.frame_bytes = d->frame_bytes;
	g.staging = d->n_threads ? (d->n_threads + 2 > E264_MAX_BUILDS ? E264_MAX_BUILDS : d->n_threads + 2) + 2 : 0;   /* staging areas wanted: one per picture being parsed + two in flight on the device; 0 = the backend's default */
	hostbufs_free_all(d);
	if (d->be->configure(d->be_ctx, &g, d->n_slots)) return ENOMEM;
	int oom = 0;
	if (d->n_threads) pthread_mutex_lock(&d->lock);
	for (int i = 0; i < E264_MAX_BUILDS; i++) {
		free(d->builds[i].mbi); memset(&d->builds[i], 0, sizeof(d->builds[i]));
		if (i < d->n_builds && !(d->builds[i].mbi = (MbInfo *)calloc((size_t)d->w_mbs * d->h_mbs, sizeof(MbInfo)))) oom = 1;
	}
	d->pb = NULL;
	memset(d->slot_users, 0, sizeof(d->slot_users)); memset(d->slot_build, 0, sizeof(d->slot_build));
	if (d->n_threads) pthread_mutex_unlock(&d->lock);
	if (oom) return ENOMEM;
	memset(d->pics, 0, sizeof(d->pics));
	for (int i = 0; i < E264_MAX_SLOTS; i++) d->pics[i].host_buf = -1;
	d->cur = -1;
	Edge264Frame *o = &d->out_fmt; memset(o, 0, sizeof(*o));
	o->bit_depth_Y = o->bit_depth_C = 8;
	o->width_Y = (int16_t)(w - s->crop[0] - s->crop[1]); o->height_Y = (int16_t)(h - s->crop[2] - s->crop[3]);
	o->width_C = o->width_Y >> 1; o->height_C = o->height_Y >> 1;
	o->stride_Y = (int16_t)d->stride_y; o->stride_C = (int16_t)d->stride_c;
	o->stride_mb = (int16_t)(s->width_mbs * 304);   /* sizeof(Edge264Macroblock) in the reference; wraps like there */
	o->frame_crop_offsets[0] = (int16_t)s->crop[2]; o->frame_crop_offsets[1] = (int16_t)s->crop[1];
	o->frame_crop_offsets[2] = (int16_t)s->crop[3]; o->frame_crop_offsets[3] = (int16_t)s->crop[0];
	d->configured = 1;
	if (!keep_numbering) { d->q_prev_ref_frame_num = -1; d->prev_ref_frame_num = -1; d->prev_poc_msb = d->prev_poc_lsb = 0; }
	return 0;
}

static inline int clamp8(int v) { return v < -128 ? -128 : v > 127 ? 127 : v; }   /* bounded reads like the reference's get_se16(-128,127) (headers.c:721-731) */

/* ------------------------------------------------------------------------------------------ */
/* slice header                                                                                 */
/* ------------------------------------------------------------------------------------------ */
static int parse_slice_header(Edge264Decoder *d, BitReader *b, int nal_unit_type, int nal_ref_idc, SliceHeader *h, const PPS **ppps) {
	const SPS *s = &d->sps;
	memset(h, 0, sizeof(*h));
	h->first_mb = br_ue_i(b, 139264);   /* bounded like the reference's get_ue32(139263) (headers.c:968); anything larger fails the first_mb >= total test */
	int st = br_ue_i(b, 10);
	if (st > 9) return EBADMSG;
	if (nal_unit_type == 5 || s->max_num_ref_frames == 0) st = 2;   /* reference headers.c:983 */
	h->slice_type = st < 5 ? st : st - 5;
	h->pps_id = br_ue_i(b, 256);
	if (h->slice_type > 2 || h->pps_id >= 4) return ENOTSUP;
	const PPS *p = &d->pps[h->pps_id];
	if (!s->valid || !p->valid) return EBADMSG;
	*ppps = p;
	h->frame_num = br_u(b, s->log2_max_frame_num);
	if (nal_unit_type == 5) h->frame_num = 0;
	h->idr_pic_id = -1;
	if (nal_unit_type == 5) h->idr_pic_id = br_ue_i(b, 65536);
	if (s->poc_type == 0) {
		h->poc_lsb = br_u(b, s->log2_max_poc_lsb);
		if (p->bottom_field_pic_order_present) h->delta_poc_bottom = br_se(b);
	} else if (s->poc_type == 1 && !s->delta_pic_order_always_zero_flag) {
		h->delta_poc[0] = br_se(b);
		if (p->bottom_field_pic_order_present) h->delta_poc[1] = br_se(b);
	}
	h->num_ref[0] = p->num_ref_idx_default[0]; h->num_ref[1] = p->num_ref_idx_default[1];
	if (h->slice_type < 2) {
		if (h->slice_type == 1) h->direct_spatial = br_u1(b);
		if (br_u1(b)) {
			h->num_ref[0] = br_ue_i(b, 32) + 1;
			if (h->slice_type == 1) h->num_ref[1] = br_ue_i(b, 32) + 1;
		}
		if (h->num_ref[0] > 16) h->num_ref[0] = h->num_ref[0] > 32 ? 0 : 16;
		if (h->num_ref[1] > 16) h->num_ref[1] = h->num_ref[1] > 32 ? 0 : 16;
		if (h->num_ref[0] == 0 || h->num_ref[1] == 0) return EBADMSG;
		for (int l = 0; l <= h->slice_type; l++) {
			if (!br_u1(b)) continue;
			int n = 0;
			for (;;) {
				unsigned op = br_ue(b);
				if (op == 3) break;
				if (op > 5 || n >= 32) return EBADMSG;
				h->mod[l][n].op = (uint8_t)op; h->mod[l][n].val = (uint32_t)br_ue_i(b, 1 << 20); n++;
			}
			h->n_mod[l] = n;
		}
		int wp = h->slice_type == 0 ? p->weighted_pred_flag : p->weighted_bipred_idc;
		if (wp == 1) {
			h->luma_log2_wd = br_ue_i(b, 8); h->chroma_log2_wd = br_ue_i(b, 8);
			if (h->luma_log2_wd > 7 || h->chroma_log2_wd > 7) return EBADMSG;
			for (int l = 0; l <= h->slice_type; l++) for (int i = 0; i < h->num_ref[l]; i++) {
				if (br_u1(b)) { h->w[l][i][0] = (int16_t)clamp8(br_se(b)); h->o[l][i][0] = (int16_t)clamp8(br_se(b)); }
				else { h->w[l][i][0] = (int16_t)(1 << h->luma_log2_wd); h->o[l][i][0] = 0; }
				if (br_u1(b)) for (int c = 1; c < 3; c++) { h->w[l][i][c] = (int16_t)clamp8(br_se(b)); h->o[l][i][c] = (int16_t)clamp8(br_se(b)); }
				else for (int c = 1; c < 3; c++) { h->w[l][i][c] = (int16_t)(1 << h->chroma_log2_wd); h->o[l][i][c] = 0; }
			}
		}
	}
	if (nal_ref_idc) {
		if (nal_unit_type == 5) { h->no_output_of_prior_pics = br_u1(b); h->long_term_reference_flag = br_u1(b); }
		else if (br_u1(b)) {
			h->adaptive_marking = 1;
			int n = 0;
			for (;;) {
				unsigned op = br_ue(b);
				if (op == 0) break;
				if (op > 6 || n >= 64) return EBADMSG;
				h->mmco[n].op = (uint8_t)op;
				if (op == 1 || op == 3) h->mmco[n].a = (uint32_t)br_ue_i(b, 1 << 20);
				if (op == 2 || op == 4 || op == 6) h->mmco[n].a = (uint32_t)br_ue_i(b, 1 << 20);
				if (op == 3) h->mmco[n].b = (uint32_t)br_ue_i(b, 1 << 20);
				n++;
			}
			h->n_mmco = n;
		}
	}
	if (p->entropy_coding_mode && h->slice_type != 2) { h->cabac_init_idc = br_ue_i(b, 3); if (h->cabac_init_idc > 2) return EBADMSG; }
	int qd = br_se(b);
	h->slice_qp = p->pic_init_qp + qd;
	if (h->slice_qp < 0 || h->slice_qp > 51) return EBADMSG;
	if (p->deblocking_filter_control_present) {
		h->deblock_idc = br_ue_i(b, 3);
		if (h->deblock_idc > 2) return EBADMSG;
		if (h->deblock_idc != 1) {
			int a = br_se(b), bb = br_se(b);
			if (a < -6 || a > 6 || bb < -6 || bb > 6) return EBADMSG;
			h->filter_offset_a = a * 2; h->filter_offset_b = bb * 2;
		}
	}
	if (b->overrun) return EBADMSG;
	return 0;
}

/* 8.2.4: reference picture lists of the slice -> slots */
static void build_ref_lists(Edge264Decoder *d, const SliceHeader *h, int lists[2][32], int nlist[2]) {
	Pic *cur = &d->pics[d->cur];
	int st[E264_MAX_SLOTS], lt[E264_MAX_SLOTS], nst = 0, nlt = 0;
	for (int l = 0; l < 2; l++) for (int i = 0; i < 32; i++) lists[l][i] = -1;
	nlist[0] = nlist[1] = 0;
	if (d->cur_idr) return;
	for (int i = 0; i < d->n_slots; i++) {
		Pic *p = &d->pics[i];
		if (i == d->cur || !p->in_use) continue;
		if (p->ref == 1) st[nst++] = i; else if (p->ref == 2) lt[nlt++] = i;
	}
	/* long-term by LongTermFrameIdx ascending */
	for (int i = 1; i < nlt; i++) for (int j = i; j > 0 && d->pics[lt[j]].long_term_idx < d->pics[lt[j - 1]].long_term_idx; j--) { int t = lt[j]; lt[j] = lt[j - 1]; lt[j - 1] = t; }
	if (h->slice_type == 0) {
		/* short-term by FrameNum descending (absolute frame numbers, monotonic) */
		for (int i = 1; i < nst; i++) for (int j = i; j > 0 && d->pics[st[j]].frame_num > d->pics[st[j - 1]].frame_num; j--) { int t = st[j]; st[j] = st[j - 1]; st[j - 1] = t; }
		int n = 0;
		for (int i = 0; i < nst; i++) lists[0][n++] = st[i];
		for (int i = 0; i < nlt; i++) lists[0][n++] = lt[i];
		nlist[0] = n;
	} else {
		int before[E264_MAX_SLOTS], after[E264_MAX_SLOTS], nb = 0, na = 0;
		for (int i = 0; i < nst; i++) { if (d->pics[st[i]].poc_top <= cur->poc_top_dec) before[nb++] = st[i]; else after[na++] = st[i]; }
		for (int i = 1; i < nb; i++) for (int j = i; j > 0 && d->pics[before[j]].poc_top > d->pics[before[j - 1]].poc_top; j--) { int t = before[j]; before[j] = before[j - 1]; before[j - 1] = t; }
		for (int i = 1; i < na; i++) for (int j = i; j > 0 && d->pics[after[j]].poc_top < d->pics[after[j - 1]].poc_top; j--) { int t = after[j]; after[j] = after[j - 1]; after[j - 1] = t; }
		int n = 0;
		for (int i = 0; i < nb; i++) lists[0][n++] = before[i];
		for (int i = 0; i < na; i++) lists[0][n++] = after[i];
		for (int i = 0; i < nlt; i++) lists[0][n++] = lt[i];
		nlist[0] = n; n = 0;
		for (int i = 0; i < na; i++) lists[1][n++] = after[i];
		for (int i = 0; i < nb; i++) lists[1][n++] = before[i];
		for (int i = 0; i < nlt; i++) lists[1][n++] = lt[i];
		nlist[1] = n;
		if (n > 1 && lists[0][0] == lists[1][0]) { int t = lists[1][0]; lists[1][0] = lists[1][1]; lists[1][1] = t; }
	}
	/* 8.2.4.3 modification */
	int mask = (1 << d->sps.log2_max_frame_num) - 1;
	for (int l = 0; l <= (h->slice_type == 1); l++) {
		int pic_num_pred = cur->frame_num;
		for (int k = 0; k < h->n_mod[l]; k++) {
			int op = h->mod[l][k].op, pic = -1;
			if (op < 2) {
				pic_num_pred = op == 0 ? pic_num_pred - ((int)h->mod[l][k].val + 1) : pic_num_pred + ((int)h->mod[l][k].val + 1);
				for (int i = 0; i < d->n_slots; i++) if (i != d->cur && d->pics[i].in_use && d->pics[i].ref == 1 && !((d->pics[i].frame_num ^ pic_num_pred) & mask)) { pic = i; break; }
			} else if (op == 2) {
				for (int i = 0; i < d->n_slots; i++) if (i != d->cur && d->pics[i].in_use && d->pics[i].ref == 2 && d->pics[i].long_term_idx == (int)h->mod[l][k].val) { pic = i; break; }
			}
			if (pic < 0) continue;
			if (k >= h->num_ref[l] || k >= 32) break;   /* more operations than list entries: the reference stops at refIdx 32 too (headers.c:817) */
			int buf = pic, c = k;
			do { int sw = lists[l][c]; lists[l][c] = buf; buf = sw; } while (++c < h->num_ref[l] && c < 32 && buf != pic);
		}
	}
}

/* ------------------------------------------------------------------------------------------ */
/* slice NAL                                                                                    */
/* ------------------------------------------------------------------------------------------ */
static int find_free_slot(Edge264Decoder *d) {
	if (d->n_threads) {   /* prefer a slot no parser is still using: taking a busy one would wait for its picture to be submitted */
		int pick = -1;
		pthread_mutex_lock(&d->lock);
		for (int i = 0; i < d->n_slots && pick < 0; i++) if (!d->pics[i].in_use && d->slot_users[i] == 0) pick = i;
		pthread_mutex_unlock(&d->lock);
		if (pick >= 0) return pick;
	}
	for (int i = 0; i < d->n_slots; i++) if (!d->pics[i].in_use) return i;
	return -1;
}

static PicBuild *build_acquire(Edge264Decoder *d, int slot);
static void slice_enqueue(Edge264Decoder *d, PicBuild *pb, SliceJob *job, int col_slot);

static int decode_slice(Edge264Decoder *d, int nal_unit_type, int nal_ref_idc, BitReader *b, SliceJob *job) {
	SliceHeader *h = &d->sh;
	const PPS *p = NULL;
	int ret = parse_slice_header(d, b, nal_unit_type, nal_ref_idc, h, &p);
	if (ret) return ret;
	const SPS *s = &d->sps;
	if (!d->configured) return EBADMSG;
	int idr = nal_unit_type == 5;
	int mask = (1 << s->log2_max_frame_num) - 1;

	/* 7.4.1.2.4 first slice of a new picture? (same tests as the reference, headers.c:1027-1050) */
	if (d->cur >= 0) {
		const SliceHeader *f = &d->first_sh;
		int newpic = h->frame_num != f->frame_num || (nal_ref_idc > 0) != (d->cur_nal_ref_idc > 0) || h->idr_pic_id != f->idr_pic_id ||
		             (s->poc_type == 0 && h->poc_lsb != f->poc_lsb) || (s->poc_type == 1 && h->delta_poc[0] != f->delta_poc[0]);
		if (newpic) { ret = finish_picture(d); if (ret) return ret; }
	}

	if (d->cur < 0) {
		if (d->outq_n > 16) {
			if (d->n_threads) {   /* the frame the application is told to fetch may still be with a parser: wait for it here, not in a polling loop */
				pthread_mutex_lock(&d->lock);
				while (!__atomic_load_n(&d->hb[d->outq[0]].submitted, __ATOMIC_ACQUIRE)) pthread_cond_wait(&d->done_cv, &d->lock);
				pthread_mutex_unlock(&d->lock);
			}
			return ENOBUFS;
		}
		int hbuf = hostbuf_acquire(d);
		if (hbuf < 0) return ENOBUFS;
		/* frame_num and POC (8.2.1), with absolute (unwrapped) frame numbers like the reference (headers.c:1059) */
		int prev = d->prev_ref_frame_num;
		int frame_num_abs = idr ? 0 : prev + 1 + ((h->frame_num - prev - 1) & mask);
		if (idr) { d->prev_ref_frame_num = -1; d->prev_poc_msb = d->prev_poc_lsb = 0; }
		int poc, poc_top;
		if (s->poc_type == 0) {
			int sh = 32 - s->log2_max_poc_lsb;
			int prev_poc = idr ? 0 : d->prev_poc_msb;   /* prev_poc_msb holds the full POC of the previous reference picture */
			int inc = (int)((unsigned)(h->poc_lsb - prev_poc) << sh) >> sh;
			poc = poc_top = prev_poc + inc;
			if (h->delta_poc_bottom < 0) poc += h->delta_poc_bottom;   /* PicOrderCnt = min(top, bottom) */
		} else if (s->poc_type == 1) {
			int abs_fn = s->num_ref_frames_in_poc_cycle > 0 ? frame_num_abs : 0;
			if (nal_ref_idc == 0 && abs_fn > 0) abs_fn--;
			int top = h->delta_poc[0] + (nal_ref_idc ? 0 : s->offset_for_non_ref_pic);
			if (abs_fn > 0) {
				int cyc = (abs_fn - 1) / s->num_ref_frames_in_poc_cycle, in = (abs_fn - 1) % s->num_ref_frames_in_poc_cycle, sum = 0, tot = 0;
				for (int i = 0; i < s->num_ref_frames_in_poc_cycle; i++) { tot += s->offset_for_ref_frame[i]; if (i <= in) sum += s->offset_for_ref_frame[i]; }
				top += cyc * tot + sum;
			}
			int bot = top + s->offset_for_top_to_bottom_field + h->delta_poc[1];
			poc = top < bot ? top : bot; poc_top = top;
		} else poc = poc_top = frame_num_abs * 2 + (nal_ref_idc != 0) - 1;

		/* 8.2.5.2 gaps in frame_num (a reference picture was lost): like the reference (headers.c:1095-1139) insert
		 * "non-existing" short-term frames for the missing numbers — they take FrameIds and list positions and push
		 * older pictures out through the sliding window, which keeps the reference indices of the surviving pictures
		 * what the encoder meant.  Their samples are undefined in the reference (never-written buffers); ours are
		 * cleared. */
		int gap_frames = 0;
		if (!idr && prev >= 0 && frame_num_abs - prev > 1) {
			int nlong = 0, nshort = 0;
			for (int i = 0; i < d->n_slots; i++) if (d->pics[i].in_use) { nlong += d->pics[i].ref == 2; nshort += d->pics[i].ref == 1; }
			int room = s->max_num_ref_frames - nlong;
			gap_frames = frame_num_abs - prev - 1 < room ? frame_num_abs - prev - 1 : room;
			if (gap_frames < 0) gap_frames = 0;
			for (; gap_frames + nshort > room && nshort > 0; nshort--) {
				int old = -1;
				for (int i = 0; i < d->n_slots; i++) if (d->pics[i].in_use && d->pics[i].ref == 1 && (old < 0 || d->pics[i].frame_num < d->pics[old].frame_num)) old = i;
				d->pics[old].ref = 0; slot_release_if_unused(d, old);
			}
			if (gap_frames > 0) threads_drain(d);   /* the fills below go straight to the device: every earlier picture must be on its way first */
			for (int fn = frame_num_abs - gap_frames; fn < frame_num_abs; fn++) {
				int sl = find_free_slot(d);
				while (sl < 0 && bump_frame(d, -1)) sl = find_free_slot(d);
				if (sl < 0) { d->next_uid++; continue; }   /* no room even after bumping: keep at least the FrameId numbering */
				Pic *np = &d->pics[sl];
				memset(np, 0, sizeof(*np));
				np->in_use = 1; np->ref = 1; np->nonexisting = 1; np->frame_num = fn; np->uid = d->next_uid++; np->host_buf = -1;
				int npoc = 0;
				if (s->poc_type == 2) npoc = fn * 2;
				else if (s->poc_type == 1 && s->num_ref_frames_in_poc_cycle > 0) {
					int tot = 0, part = 0, in = fn % s->num_ref_frames_in_poc_cycle;
					for (int i = 0; i < s->num_ref_frames_in_poc_cycle; i++) { tot += s->offset_for_ref_frame[i]; if (i < in) part += s->offset_for_ref_frame[i]; }
					npoc = (fn / s->num_ref_frames_in_poc_cycle) * tot + part;   /* as the reference computes it (headers.c:1133-1139) */
				}
				np->poc = np->poc_dec = np->poc_top = np->poc_top_dec = npoc;
				for (int i = 0; i < E264_MAX_SLOTS; i++) np->slot_uid[i] = -1;
				if (d->be->fill_slot) { if (d->n_threads) pthread_mutex_lock(&d->be_lock); d->be->fill_slot(d->be_ctx, sl, 0, 0); if (d->n_threads) pthread_mutex_unlock(&d->be_lock); }
			}
			d->prev_ref_frame_num = frame_num_abs - 1;   /* the last inserted frame is the previous reference frame now (headers.c:1131) */
		}

		int slot = find_free_slot(d);
		if (slot < 0) {   /* DPB invariant broken (stream exceeds its own limits): drop the oldest output-pending picture */
			if (!bump_frame(d, -1)) return EBADMSG;
			slot = find_free_slot(d);
			if (slot < 0) return ENOBUFS;
		}
		Pic *cp = &d->pics[slot];
		memset(cp, 0, sizeof(*cp));
		{   /* FrameId numbering: the reference keeps an absolute FrameNum across IDRs, so an IDR after other
		     * pictures looks like a frame_num gap and non-existing frames take FrameIds (headers.c:1059,1095-1139) */
			int q = d->q_prev_ref_frame_num;
			int qfn = q + 1 + (((idr ? 0 : h->frame_num) - q - 1) & mask);
			int gap = qfn - q;
			if (gap > 1 && !gap_frames) {   /* a real gap has just been given its frames (and FrameIds) above */
				int nlong = 0;
				for (int i = 0; i < d->n_slots; i++) if (d->pics[i].in_use && d->pics[i].ref == 2) nlong++;
				int room = s->max_num_ref_frames - nlong;
				d->next_uid += (gap - 1 < room ? gap - 1 : room) > 0 ? (gap - 1 < room ? gap - 1 : room) : 0;
			}
			if (gap > 1) d->q_prev_ref_frame_num = qfn - 1;
			d->q_cur_frame_num = qfn;
		}
		cp->in_use = 1; cp->frame_num = frame_num_abs; cp->poc = cp->poc_dec = poc; cp->poc_top = cp->poc_top_dec = poc_top; cp->uid = d->next_uid++; cp->host_buf = hbuf;
		d->hb[hbuf].state = 1; d->hb[hbuf].frame_id = cp->uid; d->hb[hbuf].borrowed = 0; d->hb[hbuf].submitted = 0;
		d->cur = slot; d->cur_idr = idr; d->cur_nal_ref_idc = nal_ref_idc;
		d->first_sh = *h;
		for (int i = 0; i < E264_MAX_SLOTS; i++) cp->slot_uid[i] = d->pics[i].in_use ? d->pics[i].uid : -1;
		PicBuild *pb = build_acquire(d, slot);     /* threaded mode: may wait for a parser to finish */
		E264Staging stg; memset(&stg, 0, sizeof(stg));
		int ar;
		{ PROF_BEGIN; if (d->n_threads) pthread_mutex_lock(&d->be_lock); ar = d->be->acquire_staging(d->be_ctx, slot, &stg); if (d->n_threads) pthread_mutex_unlock(&d->be_lock); PROF_END(1); }
		if (ar) {   /* give the build back as if it had never been taken (no later picture exists yet) */
			if (d->n_threads) { pthread_mutex_lock(&d->lock); d->next_seq--; d->slot_users[slot]--; d->slot_build[slot] = NULL; pb->in_use = 0; pthread_mutex_unlock(&d->lock); }
			else pb->in_use = 0;
			cp->in_use = 0; d->cur = -1; d->hb[hbuf].state = 0;
			return ENOMEM;
		}
		cp->recs = stg.recs;
		pb->recs = stg.recs; pb->coefs = stg.coefs; pb->slices = stg.slices; pb->intra_list = stg.intra_list; pb->staging = stg.handle;
		pb->coef_cap = stg.coef_capacity; pb->n_coefs = 0; pb->n_slices = 0; pb->mbs_done = 0; pb->n_intra = 0; pb->any_deblock = 0; pb->error = 0;
		pb->slot = slot; pb->host_buf = hbuf; pb->conceal_ref = -1; pb->slice_counter = 0;
		{ PROF_BEGIN; memset(pb->mbi, 0, (size_t)d->w_mbs * d->h_mbs * sizeof(MbInfo)); PROF_END(4); }
		/* records need no clearing: every macroblock of a complete picture rewrites its own (sx_one_mb) */
		d->pb = pb;

		/* IDR / MMCO5: every earlier picture leaves in output order first (reference headers.c:632, 680) */
		int has_mmco5 = 0;
		for (int k = 0; k < h->n_mmco; k++) if (h->mmco[k].op == 5) has_mmco5 = 1;
		if (idr || has_mmco5) {
			while (bump_frame(d, slot));
			if (has_mmco5) { cp->poc_top -= cp->poc; cp->poc = 0; }   /* tempPicOrderCnt subtraction (8.2.1) for output order and later pictures; frame_num restarts in apply_marking, once the picture's own lists and operations have used the real value */
		}
		/* C.4.5.3 bumping before insertion (reference headers.c:1229-1250) */
		int max_bump = s->max_num_ref_frames;
		if (!nal_ref_idc) {
			max_bump = 0;
			for (int i = 0; i < d->n_slots; i++) if (i != slot && d->pics[i].in_use && d->pics[i].needed_for_output && d->pics[i].poc_top < cp->poc_top) max_bump++;
		}
		while (dpb_fullness(d) >= s->max_dec_frame_buffering && max_bump--) bump_frame(d, slot);
		cp->needed_for_output = 1;
		if (max_bump < 0) {   /* nothing may precede it: straight to the output queue (it is returned once decoded) */
			cp->needed_for_output = 0;
			d->hb[hbuf].state = 2; d->outq[d->outq_n++] = hbuf;
		} else if (waiting_for_output(d) > s->max_num_reorder_frames) bump_frame(d, -1);
		/* POC bookkeeping for the next picture (8.2.1.1) */
		if (nal_ref_idc) { d->prev_poc_msb = has_mmco5 ? cp->poc_top : poc_top; }
		if (has_mmco5) d->prev_ref_frame_num = 0;
	}

	PicBuild *pb = d->pb;
	if (pb->n_slices >= E264_MAX_SLICES) return ENOTSUP;
	Pic *cp = &d->pics[d->cur];
	int total = d->w_mbs * d->h_mbs;
	if (h->first_mb >= total) return EBADMSG;

	/* slice record: deblocking, weights, scaling lists */
	E264SliceRec *sr = &pb->slices[pb->n_slices];
	memset(sr, 0, sizeof(*sr));
	sr->filter_offset_a = (int8_t)h->filter_offset_a; sr->filter_offset_b = (int8_t)h->filter_offset_b;
	sr->deblock_idc = (uint8_t)h->deblock_idc; sr->slice_type = (uint8_t)h->slice_type;
	merge_scaling(s, p, sr);

	SliceCtx *c = job ? &job->sc : &d->sc;
	int col_slot = -1;
	c->cabac = p->entropy_coding_mode;
	c->w_mbs = d->w_mbs; c->h_mbs = d->h_mbs;
	c->slice_type = h->slice_type; c->slice_id = ++pb->slice_counter; c->slice_idx = pb->n_slices;
	c->num_ref[0] = h->num_ref[0]; c->num_ref[1] = h->num_ref[1];
	c->direct_spatial = h->direct_spatial; c->direct_8x8_inference = s->direct_8x8_inference; c->transform_8x8_mode = p->transform_8x8_mode;
	c->qp = h->slice_qp; c->chroma_qp_offset[0] = p->chroma_qp_index_offset[0]; c->chroma_qp_offset[1] = p->chroma_qp_index_offset[1];
	c->deblock_idc = h->deblock_idc; c->cur_poc = cp->poc_dec;
	c->mbi = pb->mbi; c->recs = pb->recs; c->coefs = pb->coefs; c->n_coefs = job ? 0 : pb->n_coefs; c->coef_cap = pb->coef_cap;   /* threaded: the worker continues the pool where the previous slice ended */
	c->col_recs = NULL; c->col_slot_uid = NULL; c->error = 0; c->n_intra = 0;
	memset(c->ref_slot, -1, sizeof(c->ref_slot)); memset(c->ref_long, 0, sizeof(c->ref_long));
	for (int l = 0; l < 2; l++) for (int i = 0; i < 32; i++) { c->ref_uid[l][i] = -1; c->ref_poc[l][i] = 0; }
	if (h->slice_type < 2) {
		int lists[2][32], nl[2];
		build_ref_lists(d, h, lists, nl);
		for (int l = 0; l <= (h->slice_type == 1); l++) for (int i = 0; i < h->num_ref[l]; i++) {
			int sl = lists[l][i];
			if (sl < 0) continue;   /* missing reference: MC will read slot -1 -> treated as the current slot by the backend */
			c->ref_slot[l][i] = (int8_t)sl; c->ref_uid[l][i] = d->pics[sl].uid; c->ref_poc[l][i] = d->pics[sl].poc; c->ref_long[l][i] = d->pics[sl].ref == 2;
		}
		if (h->slice_type == 1 && lists[1][0] >= 0) {
			c->col_recs = d->pics[lists[1][0]].recs; c->col_slot_uid = d->pics[lists[1][0]].slot_uid;
			if (d->pics[lists[1][0]].nonexisting) c->col_recs = NULL; else col_slot = lists[1][0];
			if (job) {   /* the co-located picture's Pic entry may be recycled before this slice is parsed: keep a private copy of the uid table */
				memcpy(job->col_slot_uid, d->pics[lists[1][0]].slot_uid, sizeof(job->col_slot_uid)); c->col_slot_uid = job->col_slot_uid;
			}
		}
		int wp = h->slice_type == 0 ? p->weighted_pred_flag : p->weighted_bipred_idc;
		sr->wp_mode = (uint8_t)wp; sr->luma_log2_wd = (uint8_t)h->luma_log2_wd; sr->chroma_log2_wd = (uint8_t)h->chroma_log2_wd;
		if (wp == 1) for (int l = 0; l < 2; l++) for (int i = 0; i < 16; i++) for (int k = 0; k < 3; k++) { sr->wp_w[l][i][k] = h->w[l][i][k]; sr->wp_o[l][i][k] = h->o[l][i][k]; }
		if (wp == 1 && getenv("E264_DEBUG")) for (int l = 0; l <= (h->slice_type == 1); l++) for (int i = 0; i < h->num_ref[l]; i++)
			fprintf(stderr, "wp l%d ref%d logwd %d/%d w %d %d %d o %d %d %d\n", l, i, h->luma_log2_wd, h->chroma_log2_wd, h->w[l][i][0], h->w[l][i][1], h->w[l][i][2], h->o[l][i][0], h->o[l][i][1], h->o[l][i][2]);
		if (wp == 2) {   /* 8.4.2.3.1 implicit weights from POC distances */
			for (int i0 = 0; i0 < h->num_ref[0] && i0 < 16; i0++) for (int i1 = 0; i1 < h->num_ref[1] && i1 < 16; i1++) {
				int w1 = 32;
				if (c->ref_slot[0][i0] >= 0 && c->ref_slot[1][i1] >= 0 && !c->ref_long[0][i0] && !c->ref_long[1][i1]) {
					int poc0 = c->ref_poc[0][i0], poc1 = c->ref_poc[1][i1];
					int tb = cp->poc_dec - poc0, td = poc1 - poc0;
					tb = tb < -128 ? -128 : tb > 127 ? 127 : tb; td = td < -128 ? -128 : td > 127 ? 127 : td;
					if (td != 0) {
						int tx = (16384 + (td < 0 ? -td : td) / 2) / td;
						int dsf = (tb * tx + 32) >> 6; dsf = dsf < -1024 ? -1024 : dsf > 1023 ? 1023 : dsf;
						if ((dsf >> 2) >= -64 && (dsf >> 2) <= 128) w1 = dsf >> 2;
					}
				}
				sr->implicit_w1[i0][i1] = (int16_t)w1;
			}
		}
	}
	pb->n_slices++;
	if (h->deblock_idc != 1) pb->any_deblock = 1;

	/* slice data */
	c->br = *b;
	c->mbaddr = h->first_mb;
	c->cabac_init_idc_col = h->slice_type == 2 ? 0 : 1 + h->cabac_init_idc;
	if (job) {   /* threaded mode: a worker parses the slice data; the picture is closed by the next picture's first slice or a flush */
		slice_enqueue(d, pb, job, col_slot);
		return 0;
	}
	int n; { PROF_BEGIN; n = e264_parse_slice_data(c); PROF_END(0); }
	pb->n_coefs = c->n_coefs; pb->n_intra += c->n_intra;
	if (n > 0) pb->mbs_done += n;
	if (c->error) ret = c->error == 2 ? ENOMEM : EBADMSG;
	if (pb->mbs_done >= total) { int r2 = finish_picture(d); if (!ret) ret = r2; }
	return ret;
}

/* ------------------------------------------------------------------------------------------ */
/* public API                                                                                   */
/* ------------------------------------------------------------------------------------------ */
Edge264Decoder *edge264_alloc(int n_threads, Edge264LogCb log_cb, void *log_arg, int log_mbs,
                              Edge264AllocCb alloc_cb, Edge264FreeCb free_cb, void *alloc_arg) {
	(void)log_mbs;
	if (prof_on < 0) { const char *e = getenv("E264_HOST_PROFILE"); prof_on = e && atoi(e); }
	if (log_cb) return NULL;   /* like a reference build without the logs variant (edge264.c:217-220) */
	Edge264Decoder *d = (Edge264Decoder *)calloc(1, sizeof(*d));
	if (!d) return NULL;
	d->alloc_cb = (alloc_cb && free_cb) ? alloc_cb : NULL; d->free_cb = free_cb; d->alloc_arg = alloc_arg;
	d->log_cb = log_cb; d->log_arg = log_arg;
	d->cur = -1; d->pending_release = -1; d->prev_ref_frame_num = -1; d->q_prev_ref_frame_num = -1;
	d->be = e264_default_backend();
	if (!d->be || d->be->create(&d->be_ctx)) { free(d); return NULL; }
	d->n_builds = 1;
	/* n_threads as in the reference (edge264.c:223-257): 0 = parse inside edge264_decode_NAL, < 0 = one worker per CPU up
	 * to the limit, > 0 = that many workers (here they parse slice data ahead; reconstruction is the device's) */
	{ const char *e = getenv("E264_SYNC_OUTPUT"); d->sync_output = e && atoi(e) != 0; d->block_output = d->sync_output; }
	if (n_threads < 0) { long n = sysconf(_SC_NPROCESSORS_ONLN); n_threads = n < 1 ? 1 : (int)n; }
	if (n_threads > E264_MAX_THREADS) n_threads = E264_MAX_THREADS;
	if (n_threads > 0) {
		pthread_mutex_init(&d->lock, NULL); pthread_mutex_init(&d->be_lock, NULL);
		pthread_cond_init(&d->work_cv, NULL); pthread_cond_init(&d->done_cv, NULL);
		d->n_threads = n_threads;
		for (int i = 0; i < n_threads; i++) if (pthread_create(&d->threads[i], NULL, worker_main, d)) { d->n_threads = i; threads_stop(d); d->be->destroy(d->be_ctx); free(d); return NULL; }
	}
	return d;
}

void edge264_flush(Edge264Decoder *d) {
	if (!d) return;
	if (d->cur >= 0 && d->n_threads) finish_picture(d);
	threads_drain(d);
	/* drop every picture and the sequence state, keep parameter sets (reference edge264.c:261-270 clears them too) */
	for (int i = 0; i < E264_MAX_SLOTS; i++) { d->pics[i].in_use = 0; d->pics[i].host_buf = -1; }
	for (int i = 0; i < E264_MAX_HOSTBUFS; i++) if (d->hb[i].state != 3 || !d->hb[i].borrowed) d->hb[i].state = 0;
	d->outq_n = 0; d->cur = -1; d->pending_release = -1;
	d->prev_ref_frame_num = -1; d->q_prev_ref_frame_num = -1; d->prev_poc_msb = d->prev_poc_lsb = 0;
	memset(&d->sps, 0, sizeof(d->sps)); memset(d->pps, 0, sizeof(d->pps));
	d->configured = 0;
}

void edge264_free(Edge264Decoder **pd) {
	Edge264Decoder *d;
	if (!pd || !(d = *pd)) return;
	*pd = NULL;
	if (prof_on > 0) { for (int i = 0; i < 6; i++) if (prof_n[i]) fprintf(stderr, "host profile: %-18s %8.3f ms total %6ld calls %8.1f us/call\n", prof_names[i], 1e3 * prof_t[i], prof_n[i], 1e6 * prof_t[i] / prof_n[i]); memset(prof_t, 0, sizeof(prof_t)); memset(prof_n, 0, sizeof(prof_n)); }
	if (d->n_threads) { if (d->cur >= 0) finish_picture(d); threads_stop(d); }
	hostbufs_free_all(d);
	d->be->destroy(d->be_ctx);
	for (int i = 0; i < E264_MAX_BUILDS; i++) free(d->builds[i].mbi);
	free(d->rbsp); free(d);
}

static int decode_NAL_inner(Edge264Decoder *d, const uint8_t *buf, const uint8_t *end, Edge264UnrefCb unref_cb, void *unref_arg);
int edge264_decode_NAL(Edge264Decoder *d, const uint8_t *buf, const uint8_t *end, Edge264UnrefCb unref_cb, void *unref_arg) {
	if (!d || !buf) return EINVAL;
	const int r = decode_NAL_inner(d, buf, end, unref_cb, unref_arg);
	d->block_output = r == ENOBUFS || r == ENODATA || buf >= end || d->sync_output;
	return r;
}
static int decode_NAL_inner(Edge264Decoder *d, const uint8_t *buf, const uint8_t *end, Edge264UnrefCb unref_cb, void *unref_arg) {
	if (d->pending_release >= 0) { d->hb[d->pending_release].state = 0; d->pending_release = -1; }
	if (buf >= end) {
		int r = bump_all(d);
		return r ? r : ENODATA;
	}
	int nal_ref_idc = buf[0] >> 5, nal_unit_type = buf[0] & 31;
	size_t n = (size_t)(end - buf) - 1;
	if (d->n_threads && (nal_unit_type == 1 || nal_unit_type == 5)) {   /* the slice is parsed later, from its own copy of the payload */
		SliceJob *job = (SliceJob *)calloc(1, sizeof(SliceJob));
		uint8_t *rb = job ? (uint8_t *)malloc(n + 64) : NULL;
		if (!rb) { free(job); return ENOMEM; }
		size_t rn = e264_unescape(rb, buf + 1, n);
		memset(rb + rn, 0, 32);
		job->rbsp = rb;
		BitReader b; br_init(&b, rb, rn);
		int ret = decode_slice(d, nal_unit_type, nal_ref_idc, &b, job);
		if (ret) { free(rb); free(job); }      /* not queued */
		else if (unref_cb) unref_cb(0, unref_arg);
		return ret;
	}
	if (d->rbsp_cap < n + 64) { free(d->rbsp); d->rbsp_cap = n * 2 + 4096; d->rbsp = (uint8_t *)malloc(d->rbsp_cap); if (!d->rbsp) { d->rbsp_cap = 0; return ENOMEM; } }
	size_t rn = e264_unescape(d->rbsp, buf + 1, n);
	memset(d->rbsp + rn, 0, 32);
	BitReader b; br_init(&b, d->rbsp, rn);
	int ret;
	switch (nal_unit_type) {
	case 1: case 5:
		ret = decode_slice(d, nal_unit_type, nal_ref_idc, &b, NULL);
		if (ret == 0 && unref_cb) unref_cb(0, unref_arg);   /* parsing is synchronous: the NAL bytes are no longer needed */
		return ret;
	case 7: {
		SPS s; ret = parse_sps(d, &b, &s);
		if (ret == 0) {
			int same_format = d->configured && d->sps.width_mbs == s.width_mbs && d->sps.height_mbs == s.height_mbs && !memcmp(d->sps.crop, s.crop, sizeof(s.crop));
			int fits = same_format && s.max_num_ref_frames + 2 <= d->n_slots;
			if (!fits) {
				if (d->configured && bump_all(d)) return ENOBUFS;   /* frame format change: drain first (headers.c:2005-2007) */
				ret = configure_sequence(d, &s, same_format);
				if (ret) return ret;
				if (!same_format) memset(d->pps, 0, sizeof(d->pps));   /* a format change clears the decoder, PPSs included (reference clear_decoder, headers.c:133-141) */
			}
			d->sps = s;
		}
		break; }
	case 8: ret = parse_pps(d, &b); break;
	case 10: ret = 0; if (d->cur >= 0) finish_picture(d); d->prev_ref_frame_num = -1; break;   /* end of sequence */
	case 6: case 9: case 11: case 12: ret = 0; break;
	default: ret = ENOTSUP;
	}
	if (ret == 0 && unref_cb) unref_cb(0, unref_arg);
	return ret;
}

int edge264_get_frame(Edge264Decoder *d, Edge264Frame *out, int borrow) {
	if (!d || !out) return EINVAL;
	if (d->outq_n == 0) return ENOMSG;
	int hbuf = d->outq[0];
	HostBuf *hb = &d->hb[hbuf];
	if (!__atomic_load_n(&hb->submitted, __ATOMIC_ACQUIRE)) return ENOMSG;   /* queued at insertion but still being parsed */
	/* Like the reference in threaded mode (edge264.c:373: a frame whose last macroblock is not deblocked yet answers
	 * ENOMSG), a picture still on the device is "not available yet" — unless the last decode_NAL told the application
	 * to fetch frames (ENOBUFS, end of stream): then the call waits, so that no application loop ever has to poll. */
	if (!d->block_output) {
		int pr = d->be->poll(d->be_ctx, hb->ticket);
		if (pr == EAGAIN) return ENOMSG;
		if (pr) return EIO;
	}
	{ PROF_BEGIN; int wr = d->be->wait(d->be_ctx, hb->ticket); PROF_END(3); if (wr) return EIO; }
	memmove(d->outq, d->outq + 1, (size_t)(--d->outq_n) * sizeof(int));
	*out = d->out_fmt;
	int top = out->frame_crop_offsets[0], left = out->frame_crop_offsets[3];
	out->samples[0] = hb->p + top * d->stride_y + left;
	out->samples[1] = hb->p + d->plane_y + (top >> 1) * d->stride_c + (left >> 1);
	out->samples[2] = out->samples[1] + (d->stride_c >> 1);
	out->FrameId = hb->frame_id;
	out->return_arg = (void *)(uintptr_t)(hbuf + 1);
	hb->state = 3; hb->borrowed = borrow != 0;
	if (!borrow) d->pending_release = hbuf;   /* valid until the next decode_NAL, as in the reference */
	return 0;
}

void edge264_return_frame(Edge264Decoder *d, void *return_arg) {
	if (!d) return;
	int i = (int)(uintptr_t)return_arg - 1;
	if (i >= 0 && i < E264_MAX_HOSTBUFS && d->hb[i].state == 3) { d->hb[i].state = 0; d->hb[i].borrowed = 0; if (d->pending_release == i) d->pending_release = -1; }
}
