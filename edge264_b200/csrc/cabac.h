/* cabac.h — CABAC arithmetic decoding engine (9.3.3.2) for the product parser and the matching
 * encoding engine (9.3.4.2) for the synthetic-stream generator.
 * Plays the role of the reference's get_ae/get_bypass/cabac_start/cabac_terminate/cabac_init
 * (reference: edge264_bitstream.c:256-347) with a different organisation: 9-bit range, a 64-bit
 * window `val` = codIOffset followed by `nbits` look-ahead bits, refilled 32 bits at a time,
 * context state = pStateIdx*2 + valMPS. */
#ifndef E264B_CABAC_H
#define E264B_CABAC_H
#include <stdint.h>
#include <string.h>
#include "h264_tables.h"
#include "h264_vlc_tables.h"
#include "bits.h"

typedef struct CabacDec {
	uint64_t val;
	uint32_t range;
	int nbits;
	const uint8_t *p, *start, *end;   /* buffer must carry >= 16 readable slack bytes after `end` */
	uint8_t state[1024];
} CabacDec;

static uint8_t cabac_next_mps[128], cabac_next_lps[128];
static int cabac_tables_ready;
static void cabac_build_tables(void) {
	if (cabac_tables_ready) return;
	for (int s = 0; s < 64; s++) for (int m = 0; m < 2; m++) {
		cabac_next_mps[s * 2 + m] = (uint8_t)((s < 62 ? s + 1 : s) * 2 + m);
		cabac_next_lps[s * 2 + m] = (uint8_t)(h264_trans_lps[s] * 2 + (s == 0 ? !m : m));
	}
	cabac_tables_ready = 1;
}

/* 9.3.1.1 context initialisation; col = 0 for I slices, else 1 + cabac_init_idc */
static void cabac_init_states(uint8_t *state, int col, int slice_qp) {
	int qp = slice_qp < 0 ? 0 : slice_qp > 51 ? 51 : slice_qp;
	for (int i = 0; i < 1024; i++) {
		int m = h264_cabac_mn[i][col][0], n = h264_cabac_mn[i][col][1];
		int pre = ((m * qp) >> 4) + n;
		pre = pre < 1 ? 1 : pre > 126 ? 126 : pre;
		state[i] = pre <= 63 ? (uint8_t)((63 - pre) << 1) : (uint8_t)(((pre - 64) << 1) | 1);
	}
}

static inline void cabac_refill(CabacDec *c) {
	if (c->nbits < 16) {
		uint32_t w; memcpy(&w, c->p, 4); c->p += 4;
		c->val = (c->val << 32) | __builtin_bswap32(w);
		c->nbits += 32;
	}
}
/* 9.3.1.2: p must be byte aligned */
static inline void cabac_dec_start(CabacDec *c, const uint8_t *p, const uint8_t *end) {
	c->start = c->p = p; c->end = end;
	c->val = 0; c->nbits = -9; c->range = 510;
	cabac_refill(c);
}
static inline int cabac_bin(CabacDec *c, int ctx) {
	cabac_refill(c);
	uint32_t s = c->state[ctx];
	uint32_t lps = h264_range_lps[s >> 1][(c->range >> 6) & 3];
	uint32_t rmps = c->range - lps;
	uint64_t scaled = (uint64_t)rmps << c->nbits;
	if (c->val < scaled) {
		c->state[ctx] = cabac_next_mps[s];
		int sh = rmps < 256;
		c->range = rmps << sh; c->nbits -= sh;
		return s & 1;
	}
	c->val -= scaled;
	c->state[ctx] = cabac_next_lps[s];
	int n = __builtin_clz(lps) - 23;
	c->range = lps << n; c->nbits -= n;
	return (s & 1) ^ 1;
}
static inline int cabac_bypass(CabacDec *c) {
	cabac_refill(c);
	c->nbits--;
	uint64_t scaled = (uint64_t)c->range << c->nbits;
	if (c->val >= scaled) { c->val -= scaled; return 1; }
	return 0;
}
static inline int cabac_terminate(CabacDec *c) {
	cabac_refill(c);
	c->range -= 2;
	uint64_t scaled = (uint64_t)c->range << c->nbits;
	if (c->val >= scaled) return 1;
	if (c->range < 256) { c->range <<= 1; c->nbits--; }
	return 0;
}
/* after cabac_terminate()==1: byte position of the first byte after the (aligned) end of the
 * arithmetic code word, i.e. where pcm samples start (7.3.5 pcm_alignment_zero_bit) */
static inline const uint8_t *cabac_dec_aligned_pos(const CabacDec *c) {
	int64_t bits = (int64_t)(c->p - c->start) * 8 - c->nbits;   /* consumed into codIOffset, incl. the stop bit */
	return c->start + ((bits + 7) >> 3);
}
static inline int cabac_dec_overrun(const CabacDec *c) {
	return (int64_t)(c->p - c->end) * 8 - c->nbits > 16;  /* consumed clearly past the end of the RBSP */
}

/* ---------------- encoder (generator only) ---------------- */
typedef struct CabacEnc {
	uint32_t low, range; int outstanding, first;
	BitWriter *w;
	uint8_t state[1024];
} CabacEnc;
static inline void cabac_enc_start(CabacEnc *e, BitWriter *w) { e->w = w; e->low = 0; e->range = 510; e->outstanding = 0; e->first = 1; }
static inline void cabac_enc_put(CabacEnc *e, int b) {
	if (e->first) e->first = 0; else bw_u(e->w, 1, b);
	while (e->outstanding > 0) { bw_u(e->w, 1, 1 - b); e->outstanding--; }
}
static inline void cabac_enc_renorm(CabacEnc *e) {
	while (e->range < 256) {
		if (e->low < 256) cabac_enc_put(e, 0);
		else if (e->low >= 512) { e->low -= 512; cabac_enc_put(e, 1); }
		else { e->low -= 256; e->outstanding++; }
		e->range <<= 1; e->low <<= 1;
	}
}
static inline int cabac_enc_bin(CabacEnc *e, int ctx, int bin) {
	uint32_t s = e->state[ctx];
	uint32_t lps = h264_range_lps[s >> 1][(e->range >> 6) & 3];
	e->range -= lps;
	if ((int)(s & 1) != bin) { e->low += e->range; e->range = lps; e->state[ctx] = cabac_next_lps[s]; }
	else e->state[ctx] = cabac_next_mps[s];
	cabac_enc_renorm(e);
	return bin;
}
static inline int cabac_enc_bypass(CabacEnc *e, int bin) {
	e->low <<= 1;
	if (bin) e->low += e->range;
	if (e->low >= 1024) { cabac_enc_put(e, 1); e->low -= 1024; }
	else if (e->low < 512) cabac_enc_put(e, 0);
	else { e->low -= 512; e->outstanding++; }
	return bin;
}
static inline int cabac_enc_terminate(CabacEnc *e, int bin) {
	e->range -= 2;
	if (bin) {
		e->low += e->range;
		e->range = 2;
		cabac_enc_renorm(e);
		cabac_enc_put(e, (e->low >> 9) & 1);
		bw_u(e->w, 2, ((e->low >> 7) & 3) | 1);   /* last written bit doubles as rbsp_stop_one_bit */
	} else cabac_enc_renorm(e);
	return bin;
}
#endif
