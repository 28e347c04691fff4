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

/* Decoder registers: codIOffset sits at a FIXED position (bits 62..54 of `val`, one bit of headroom for
 * bypass doubling) followed by `avail` valid look-ahead bits, so the MPS/LPS comparison needs only a
 * constant shift; renormalisation is one count-leading-zeros and the LPS/MPS selection is branch-free. */
/* context states are 16-bit on purpose: byte stores may alias anything, which would force the compiler to
 * reload range/val/avail after every bin; int16 stores cannot alias them (strict aliasing) */
typedef int16_t CabacState;

typedef struct CabacDec {
	uint64_t val;
	uint32_t range;
	int avail;
	const uint8_t *p, *start, *end;   /* buffer must carry >= 16 readable slack bytes after `end` */
	CabacState state[1024];
} CabacDec;
#define CABAC_POS 54

static uint8_t cabac_next_mps[128], cabac_next_lps[128];
static uint8_t cabac_trans[256];      /* [0..127] state after an MPS, [128..255] after an LPS */
static uint8_t cabac_lps4[4 * 128];   /* rangeTabLPS indexed by qCodIRangeIdx * 128 + packed state */
/* per packed state: four 16-bit entries (one per qCodIRangeIdx) = rangeLPS | renormalisation shift of rangeLPS << 8.
 * Indexed by the STATE only, so the load does not wait for the previous bin's range; the range picks its entry
 * with one variable shift. */
static uint64_t cabac_lpsw[128];
static int cabac_tables_ready;
static void cabac_build_tables(void) {
	if (cabac_tables_ready) return;
	for (int s = 0; s < 64; s++) for (int m = 0; m < 2; m++) {
		cabac_next_mps[s * 2 + m] = (uint8_t)((s < 62 ? s + 1 : s) * 2 + m);
		cabac_next_lps[s * 2 + m] = (uint8_t)(h264_trans_lps[s] * 2 + (s == 0 ? !m : m));
		cabac_trans[s * 2 + m] = cabac_next_mps[s * 2 + m];
		cabac_trans[128 + s * 2 + m] = cabac_next_lps[s * 2 + m];
		for (int q = 0; q < 4; q++) cabac_lps4[q * 128 + s * 2 + m] = h264_range_lps[s][q];
		uint64_t w = 0;
		for (int q = 0; q < 4; q++) w |= (uint64_t)(h264_range_lps[s][q] | ((__builtin_clz((unsigned)h264_range_lps[s][q]) - 23) << 8)) << (16 * q);
		cabac_lpsw[s * 2 + m] = w;
	}
	cabac_tables_ready = 1;
}

/* 9.3.1.1 context initialisation; col = 0 for I slices, else 1 + cabac_init_idc */
static void cabac_init_states(CabacState *state, int col, int slice_qp) {
	int qp = slice_qp < 0 ? 0 : slice_qp > 51 ? 51 : slice_qp;
	for (int i = 0; i < 1024; i++) {
		int m = h264_cabac_mn[i][col][0], n = h264_cabac_mn[i][col][1];
		int pre = ((m * qp) >> 4) + n;
		pre = pre < 1 ? 1 : pre > 126 ? 126 : pre;
		state[i] = pre <= 63 ? (CabacState)((63 - pre) << 1) : (CabacState)(((pre - 64) << 1) | 1);
	}
}

/* the four hot fields, kept in locals by the residual parser */
typedef struct CabacRegs { uint64_t val; uint32_t range; int avail; const uint8_t *p; } CabacRegs;
static inline CabacRegs cabac_regs_load(const CabacDec *c) { CabacRegs r = {c->val, c->range, c->avail, c->p}; return r; }
static inline void cabac_regs_store(CabacDec *c, const CabacRegs *r) { c->val = r->val; c->range = r->range; c->avail = r->avail; c->p = r->p; }

static inline void cabac_r_refill(CabacRegs *r) {
	if (__builtin_expect(r->avail <= CABAC_POS - 32, 0)) {
		uint32_t w; memcpy(&w, r->p, 4); r->p += 4;
		r->val |= (uint64_t)__builtin_bswap32(w) << (CABAC_POS - 32 - r->avail);
		r->avail += 32;
	}
}
/* One context-coded bin.  The serial dependency between bins runs through `range` (and `val`); it is kept short:
 * the table word depends on the state only, the renormalised LPS range comes from the table (no count-leading-zeros
 * in the chain), the MPS range needs a shift of 0 or 1, and the two outcomes are selected with conditional moves. */
static inline int cabac_r_bin(CabacRegs *r, CabacState *state, int ctx) {
	cabac_r_refill(r);
	const uint32_t s = state[ctx];
	const uint64_t w = cabac_lpsw[s];
	const uint32_t range = r->range;
	const uint32_t e = (uint32_t)(w >> ((range >> 2) & 0x30));
	const uint32_t lps = e & 0xff, nl = (e >> 8) & 7;
	const uint32_t rmps = range - lps;
	const uint64_t scaled = (uint64_t)rmps << CABAC_POS;
	const uint32_t nm = (rmps >> 8) ^ 1;                 /* rmps is in [128, 510]: shift by one iff below 256 */
	const uint32_t rm = rmps << nm, rl = lps << nl;
	uint64_t val = r->val; const uint64_t vl = val - scaled;
	uint32_t n = nm, range2 = rm, is_lps;
#if defined(__x86_64__)
	/* compilers turn the three selections into ONE branch, and bins at this bit rate are close to unpredictable
	 * (about 0.75 bit of information each): conditional moves keep the cost flat */
	uint8_t f;
	__asm__("cmp %[sc], %[v]\n\tcmovae %[rl], %[rg]\n\tcmovae %[nl], %[n]\n\tcmovae %[vl], %[v]\n\tsetae %[f]"
		: [rg] "+r"(range2), [n] "+r"(n), [v] "+r"(val), [f] "=q"(f)
		: [sc] "r"(scaled), [rl] "r"(rl), [nl] "r"(nl), [vl] "r"(vl) : "cc");
	is_lps = f;
#else
	const uint64_t m = (uint64_t)0 - (uint64_t)(val >= scaled);
	is_lps = (uint32_t)m & 1; n ^= (n ^ nl) & (uint32_t)m; range2 ^= (range2 ^ rl) & (uint32_t)m; val -= scaled & m;
#endif
	r->range = range2;
	r->val = val << n;
	r->avail -= n;
	state[ctx] = cabac_trans[s + (is_lps << 7)];
	return (int)((s & 1) ^ is_lps);
}
static inline int cabac_r_bypass(CabacRegs *r) {
	cabac_r_refill(r);
	r->val <<= 1; r->avail--;
	uint64_t scaled = (uint64_t)r->range << CABAC_POS;
	uint64_t m = (uint64_t)0 - (uint64_t)(r->val >= scaled);
	r->val -= scaled & m;
	return (int)(m & 1);
}

/* 9.3.1.2: p must be byte aligned */
static inline void cabac_dec_start(CabacDec *c, const uint8_t *p, const uint8_t *end) {
	c->start = c->p = p; c->end = end;
	c->val = 0; c->avail = -9; c->range = 510;
	/* first 32 bits: 9 into codIOffset, 23 look-ahead */
	uint32_t w; memcpy(&w, c->p, 4); c->p += 4;
	c->val = (uint64_t)__builtin_bswap32(w) << (CABAC_POS + 9 - 32);
	c->avail = 23;
}
static inline int cabac_bin(CabacDec *c, int ctx) {
	CabacRegs r = cabac_regs_load(c);
	int b = cabac_r_bin(&r, c->state, ctx);
	cabac_regs_store(c, &r);
	return b;
}
static inline int cabac_bypass(CabacDec *c) {
	CabacRegs r = cabac_regs_load(c);
	int b = cabac_r_bypass(&r);
	cabac_regs_store(c, &r);
	return b;
}
static inline int cabac_terminate(CabacDec *c) {
	CabacRegs r = cabac_regs_load(c);
	cabac_r_refill(&r);
	r.range -= 2;
	uint64_t scaled = (uint64_t)r.range << CABAC_POS;
	int bin = r.val >= scaled;
	if (!bin && r.range < 256) { r.range <<= 1; r.val <<= 1; r.avail--; }
	cabac_regs_store(c, &r);
	return bin;
}
/* after cabac_terminate()==1: byte position of the first byte after the (aligned) end of the
 * arithmetic code word, i.e. where pcm samples start (7.3.5 pcm_alignment_zero_bit) */
static inline const uint8_t *cabac_dec_aligned_pos(const CabacDec *c) {
	int64_t bits = (int64_t)(c->p - c->start) * 8 - c->avail;   /* consumed into codIOffset, incl. the stop bit */
	return c->start + ((bits + 7) >> 3);
}
static inline int cabac_dec_overrun(const CabacDec *c) {
	return (int64_t)(c->p - c->end) * 8 - c->avail > 16;  /* consumed clearly past the end of the RBSP */
}

/* ---------------- encoder (generator only) ---------------- */
typedef struct CabacEnc {
	uint32_t low, range; int outstanding, first;
	BitWriter *w;
	CabacState state[1024];
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
