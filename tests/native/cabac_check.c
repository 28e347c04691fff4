/* The product's CABAC bin decoder (edge264_b200/csrc/cabac.h: table word per state, pre-normalised LPS range, conditional
 * moves) against a literal restatement of 9.3.3.2.1 / 9.3.3.2.3 (DecodeDecision / DecodeBypass with codIRange, codIOffset,
 * RenormD bit by bit; reference counterpart edge264_bitstream.c:256-347) on bins produced by the encoder of 9.3.4.2:
 * random contexts with skewed and near-uniform probabilities, bypass bins mixed in, several slice QPs / init columns. */
#include <stdio.h>
#include <stdlib.h>
#include "../../edge264_b200/csrc/cabac.h"

typedef struct { const uint8_t *p; size_t nbits, pos; unsigned range, offset; uint8_t pstate[1024], mps[1024]; } SpecDec;
static unsigned spec_bit(SpecDec *d) { unsigned b = d->pos < d->nbits ? (d->p[d->pos >> 3] >> (7 - (d->pos & 7))) & 1 : 0; d->pos++; return b; }
static void spec_start(SpecDec *d, const uint8_t *p, size_t n) { d->p = p; d->nbits = n * 8; d->pos = 0; d->range = 510; d->offset = 0; for (int i = 0; i < 9; i++) d->offset = d->offset << 1 | spec_bit(d); }
static int spec_decision(SpecDec *d, int ctx) {
	unsigned q = (d->range >> 6) & 3, lps = h264_range_lps[d->pstate[ctx]][q];
	int bin;
	d->range -= lps;
	if (d->offset >= d->range) { bin = !d->mps[ctx]; d->offset -= d->range; d->range = lps; if (d->pstate[ctx] == 0) d->mps[ctx] ^= 1; d->pstate[ctx] = h264_trans_lps[d->pstate[ctx]]; }
	else { bin = d->mps[ctx]; if (d->pstate[ctx] < 62) d->pstate[ctx]++; }
	while (d->range < 256) { d->range <<= 1; d->offset = d->offset << 1 | spec_bit(d); }
	return bin;
}
static int spec_bypass(SpecDec *d) { d->offset = d->offset << 1 | spec_bit(d); if (d->offset >= d->range) { d->offset -= d->range; return 1; } return 0; }

static uint64_t rng = 0x9E3779B97F4A7C15ull;
static unsigned rnd(void) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (unsigned)(rng >> 33); }

int main(void) {
	cabac_build_tables();
	long total = 0, mism = 0;
	for (int trial = 0; trial < 64; trial++) {
		const int qp = (int)(rnd() % 52), col = (int)(rnd() % 4), n = 20000 + (int)(rnd() % 20000);
		int *ctxs = malloc(sizeof(int) * n), *bins = malloc(sizeof(int) * n);
		BitWriter w; bw_init(&w, 1 << 16);
		CabacEnc e; cabac_enc_start(&e, &w); cabac_init_states(e.state, col, qp);
		for (int i = 0; i < n; i++) {
			const unsigned r = rnd();
			ctxs[i] = (r & 15) == 0 ? -1 : (int)((r >> 4) % (trial & 1 ? 1024 : 12));        /* -1: bypass; few contexts = long adaptation runs */
			const int skew = trial % 4;                                                         /* 0: uniform bins, 3: strongly skewed */
			bins[i] = skew == 0 ? (int)(rnd() & 1) : (int)(rnd() % (2u << skew) == 0);
			if (ctxs[i] < 0) cabac_enc_bypass(&e, bins[i]); else cabac_enc_bin(&e, ctxs[i], bins[i]);
		}
		cabac_enc_terminate(&e, 1);
		while (w.pos & 7) bw_u(&w, 1, 0);
		size_t bytes = w.pos >> 3;
		uint8_t *buf = (uint8_t *)calloc(bytes + 64, 1); memcpy(buf, w.buf, bytes);
		CabacDec d; cabac_init_states(d.state, col, qp); cabac_dec_start(&d, buf, buf + bytes);
		SpecDec s; for (int i = 0; i < 1024; i++) { s.pstate[i] = (uint8_t)(d.state[i] >> 1); s.mps[i] = (uint8_t)(d.state[i] & 1); } spec_start(&s, buf, bytes + 8);
		CabacRegs r = cabac_regs_load(&d);
		for (int i = 0; i < n; i++) {
			int a, b;
			if (ctxs[i] < 0) { a = cabac_r_bypass(&r); b = spec_bypass(&s); }
			else { a = cabac_r_bin(&r, d.state, ctxs[i]); b = spec_decision(&s, ctxs[i]); }
			total++;
			if (a != bins[i] || b != bins[i]) { if (mism < 5) printf("trial %d bin %d ctx %d: product %d spec %d encoded %d\n", trial, i, ctxs[i], a, b, bins[i]); mism++; }
			if (ctxs[i] >= 0 && ((d.state[ctxs[i]] & 0xff) != ((s.pstate[ctxs[i]] << 1) | s.mps[ctxs[i]]))) { if (mism < 5) printf("trial %d bin %d: context state differs\n", trial, i); mism++; }
		}
		free(ctxs); free(bins); free(buf); free(w.buf);
	}
	printf("%ld bins, %ld mismatches\n", total, mism);
	return mism != 0;
}
