/* bits.h — RBSP bit reader (product) and bit writer (stream generator / tests only).
 * The reader works on an UNESCAPED copy of the NAL payload (emulation-prevention bytes removed by
 * e264_unescape), unlike the reference which unescapes on the fly in its cache refill
 * (reference: edge264_bitstream.c:13-101).  Exp-Golomb per 9.1. */
#ifndef E264B_BITS_H
#define E264B_BITS_H
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

typedef struct BitReader {
	const uint8_t *buf;
	size_t size;      /* bytes */
	size_t pos;       /* bit position */
	int overrun;
} BitReader;

/* Remove emulation_prevention_three_bytes; returns RBSP size.  dst must hold `n` bytes (+8 slack). */
static inline size_t e264_unescape(uint8_t *dst, const uint8_t *src, size_t n) {
	size_t o = 0, i = 0;
	while (i < n) {
		const uint8_t *z = (const uint8_t *)memchr(src + i, 3, n - i);
		size_t stop = z ? (size_t)(z - src) : n;
		memcpy(dst + o, src + i, stop - i); o += stop - i;
		if (!z) break;
		if (!(stop >= 2 && src[stop - 1] == 0 && src[stop - 2] == 0)) dst[o++] = 3;
		i = stop + 1;
	}
	return o;
}

static inline void br_init(BitReader *b, const uint8_t *buf, size_t size) { b->buf = buf; b->size = size; b->pos = 0; b->overrun = 0; }

static inline uint32_t br_peek32(BitReader *b) {   /* next 32 bits, MSB first, zero beyond the end */
	size_t byte = b->pos >> 3; int sh = b->pos & 7;
	uint64_t v = 0;
	if (byte + 8 <= b->size + 8) {   /* buffers carry 8 bytes of zero slack */
		uint64_t t; memcpy(&t, b->buf + byte, 8); v = __builtin_bswap64(t);
	}
	return (uint32_t)((v << sh) >> 32);
}
static inline uint32_t br_u(BitReader *b, int n) {   /* n in 0..32 */
	if (n == 0) return 0;
	uint32_t v = br_peek32(b) >> (32 - n);
	b->pos += n;
	if (b->pos > b->size * 8) b->overrun = 1;
	return v;
}
static inline uint32_t br_u1(BitReader *b) { return br_u(b, 1); }
static inline uint32_t br_ue(BitReader *b) {
	uint32_t p = br_peek32(b);
	if (p == 0) { b->pos += 32; b->overrun = 1; return 0; }
	int lz = __builtin_clz(p);
	if (lz <= 15) {
		b->pos += 2 * lz + 1;
		if (b->pos > b->size * 8) b->overrun = 1;
		return (p >> (31 - 2 * lz)) - 1;
	}
	b->pos += lz + 1;
	return (uint32_t)(((uint64_t)1 << lz) - 1 + br_u(b, lz));
}
/* bounded ue(v) for header fields that land in an int: values above `cap` (codes up to 2^32-2 exist) are clamped to
 * `cap`, which every caller chooses just above the field's legal range so that its range check fails — the
 * reference bounds every such read (get_ue16 / get_ue32 with a maximum, edge264_bitstream.c:150-203) */
static inline int br_ue_i(BitReader *b, int cap) { uint32_t v = br_ue(b); return v > (uint32_t)cap ? cap : (int)v; }
static inline int32_t br_se(BitReader *b) {
	uint32_t k = br_ue(b);
	return (k & 1) ? (int32_t)((k + 1) >> 1) : -(int32_t)(k >> 1);
}
static inline int br_bits_left(const BitReader *b) { return (int)((int64_t)b->size * 8 - (int64_t)b->pos); }
/* more_rbsp_data(): true if there is something before the rbsp_stop_one_bit */
static inline int br_more_rbsp_data(const BitReader *b) {
	int64_t last = (int64_t)b->size - 1;
	while (last >= 0 && b->buf[last] == 0) last--;
	if (last < 0) return 0;
	int tz = __builtin_ctz(b->buf[last]);
	int64_t stop_bit_pos = last * 8 + (7 - tz);
	return (int64_t)b->pos < stop_bit_pos;
}

/* ---------- writer (generator only) ---------- */
typedef struct BitWriter {
	uint8_t *buf; size_t cap; size_t pos; /* bit position */
} BitWriter;
static inline void bw_init(BitWriter *w, size_t cap) { w->buf = (uint8_t *)calloc(cap + 16, 1); w->cap = cap; w->pos = 0; }
static inline void bw_grow(BitWriter *w, size_t need_bits) {
	if ((w->pos + need_bits + 64) / 8 >= w->cap) {
		size_t ncap = w->cap * 2 + need_bits / 8 + 64;
		w->buf = (uint8_t *)realloc(w->buf, ncap + 16); memset(w->buf + w->cap, 0, ncap + 16 - w->cap); w->cap = ncap;
	}
}
static inline void bw_u(BitWriter *w, int n, uint32_t v) {
	bw_grow(w, n);
	for (int i = n - 1; i >= 0; i--) { if ((v >> i) & 1) w->buf[w->pos >> 3] |= 0x80 >> (w->pos & 7); w->pos++; }
}
static inline void bw_ue(BitWriter *w, uint32_t v) {
	uint64_t x = (uint64_t)v + 1; int len = 63 - __builtin_clzll(x);
	bw_u(w, len, 0); bw_u(w, 1, 1);
	if (len) bw_u(w, len, (uint32_t)(x & ((1ull << len) - 1)));
}
static inline void bw_se(BitWriter *w, int32_t v) { bw_ue(w, v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * (int64_t)v)); }
static inline void bw_trailing(BitWriter *w) { bw_u(w, 1, 1); while (w->pos & 7) bw_u(w, 1, 0); }
static inline void bw_align_zero(BitWriter *w) { while (w->pos & 7) bw_u(w, 1, 0); }
static inline void bw_bytes(BitWriter *w, const uint8_t *p, size_t n) { bw_grow(w, n * 8); memcpy(w->buf + (w->pos >> 3), p, n); w->pos += n * 8; }

/* append start code + NAL header + escaped payload to `out` (grown with realloc) */
typedef struct ByteBuf { uint8_t *p; size_t n, cap; } ByteBuf;
static inline void bb_put(ByteBuf *b, uint8_t v) {
	if (b->n == b->cap) { b->cap = b->cap * 2 + 4096; b->p = (uint8_t *)realloc(b->p, b->cap); }
	b->p[b->n++] = v;
}
static inline void e264_emit_nal(ByteBuf *out, int nal_ref_idc, int nal_unit_type, const uint8_t *rbsp, size_t n) {
	bb_put(out, 0); bb_put(out, 0); bb_put(out, 0); bb_put(out, 1);
	bb_put(out, (uint8_t)((nal_ref_idc << 5) | nal_unit_type));
	int zeros = 0;
	for (size_t i = 0; i < n; i++) {
		if (zeros >= 2 && rbsp[i] <= 3) { bb_put(out, 3); zeros = 0; }
		bb_put(out, rbsp[i]);
		zeros = rbsp[i] == 0 ? zeros + 1 : 0;
	}
	if (n && rbsp[n - 1] == 0) bb_put(out, 3);   /* 7.4.1: a NAL unit must not end in 0x00 (cabac_zero_words case) */
}
#endif
