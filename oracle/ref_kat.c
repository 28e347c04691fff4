/* oracle/ref_kat.c — TEST INFRASTRUCTURE (build container only).  Includes the reference's intra and inter
 * sources the way its own edge264_check.c:20-22 does, and
 *   ref_kat fuzz     differential fuzz of the reference's decode_inter_luma / decode_inter_chroma against
 *                    the plain-C restatement (oracle/port_recon.c), all 48 luma modes, random + extreme sources
 *   ref_kat dump     prints the known-answer inputs/outputs of the reference's KAT setup as JSON
 *                    (tests/golden/kat_*.json are generated from this)
 * Nothing from the reference is copied: the sources are compiled from $(REF)/src at build time. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "edge264_internal.h"
#include "edge264_intra.c"
#include "edge264_inter.c"

typedef struct { const uint8_t *p; int stride, w, h; } Plane;
int port_mc_luma_sample(const Plane *r, int x, int y, int fx, int fy);
int port_mc_chroma_sample(const Plane *r, int x, int y, int fx, int fy);
void port_intra4x4(uint8_t *p, int stride, int imode);
void port_intra8x8(uint8_t *p, int stride, int imode);
void port_intra16x16(uint8_t *p, int stride, int imode);
void port_intra_chroma(uint8_t *p, int stride, int imode);

static uint64_t rs = 88172645463325252ull;
static unsigned rnd(void) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (unsigned)(rs >> 11); }

static int fuzz_inter(void) {
	static uint8_t src[64 * 64] __attribute__((aligned(16))), dst[16 * 16] __attribute__((aligned(16)));
	int bad = 0;
	for (int trial = 0; trial < 3000 && bad < 10; trial++) {
		int kind = trial % 4;
		int px = 1 + rnd() % 3, py = 1 + rnd() % 3, ph = rnd();
		for (int i = 0; i < 64 * 64; i++) src[i] = kind == 0 ? rnd() : kind == 1 ? ((rnd() & 1) ? 255 : 0) : kind == 2 ? (uint8_t)((((i % 64) / px + (i / 64) / py + ph) & 1) ? 255 - rnd() % 3 : rnd() % 3) : ((rnd() % 5) ? 255 : rnd());
		for (int mode = 0; mode < 48; mode++) {
			int w = 4 << (mode >> 4), fx = mode & 3, fy = (mode >> 2) & 3;
			for (int hh = 0; hh < 3; hh++) {
				int h = 4 << hh;
				if ((w == 4 && h == 16) || (w == 16 && h == 4)) continue;
				memset(dst, 0, sizeof(dst));
				decode_inter_luma(mode, h, 64, src + 20 * 64 + 20, 16, dst, (i8x16){0, 1});
				Plane P = {src, 64, 64, 64};
				for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
					int v = port_mc_luma_sample(&P, 20 + x, 20 + y, fx, fy);
					if (v != dst[y * 16 + x] && bad < 10) { printf("LUMA MISMATCH trial %d mode %d (w%d h%d fx%d fy%d) at (%d,%d): ref %d port %d\n", trial, mode, w, h, fx, fy, x, y, dst[y * 16 + x], v); bad++; }
				}
			}
		}
	}
	printf("inter luma fuzz: %s\n", bad ? "MISMATCHES" : "ok");
	return bad;
}

/* Known-answer dump: the inputs of the reference's own KATs (edge264_check.c:173-180 border ramp, :286-290
 * luma source) run through the reference's functions; one JSON object on stdout. */
static void dump_block(const char *name, int mode, const uint8_t *p, int stride, int w, int h, int last) {
	printf("  {\"fn\": \"%s\", \"mode\": %d, \"w\": %d, \"h\": %d, \"out\": [", name, mode, w, h);
	for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) printf("%d%s", p[y * stride + x], (y == h - 1 && x == w - 1) ? "" : ",");
	printf("]}%s\n", last ? "" : ",");
}
static int dump_kat(void) {
	uint8_t buf[32 * 18] __attribute__((aligned(16)));
	uint8_t *p = buf + 80;
	i16x8 clip = {255, 255, 255, 255, 255, 255, 255, 255};
	printf("{\"intra\": [\n");
#define RESET_BORDER() do { memset(buf, 0, sizeof(buf)); for (int x = -1; x < 16; x++) { p[x - 32] = 194 + x * 4; p[x - 64] = 198 + x * 4; } for (int y = 0; y < 16; y++) p[y * 32 - 1] = 186 - y * 4; } while (0)
	for (int m = 0; m <= I4x4_HU_8; m++) { RESET_BORDER(); decode_intra4x4(p, 32, m, clip); dump_block("intra4x4", m, p, 32, 4, 4, 0); }
	for (int m = 0; m <= I8x8_HU_D_8; m++) { RESET_BORDER(); decode_intra8x8(p, 32, m, clip); dump_block("intra8x8", m, p, 32, 8, 8, 0); }
	for (int m = 0; m <= I16x16_P_8; m++) { RESET_BORDER(); decode_intra16x16(p, 32, m, clip); dump_block("intra16x16", m, p, 32, 16, 16, 0); }
	for (int m = 0; m <= IC8x8_P_8; m++) { RESET_BORDER(); decode_intraChroma(p, 32, m, clip); dump_block("intraChroma", m, p, 32, 8, 16, m == IC8x8_P_8); }
	printf("], \"inter_luma\": [\n");
	uint8_t src[444], dst[256] __attribute__((aligned(16)));
	for (int i = 0; i < 441; i++) src[i] = (uint8_t)(i * 37);
	for (int mode = 0; mode < 48; mode++) {
		memset(dst, 0, sizeof(dst));
		int h = mode < 16 ? 8 : 16, w = 4 << (mode >> 4);
		decode_inter_luma(mode, h, 21, src + 44, 16, dst, (i8x16){0, 1});
		dump_block("inter_luma", mode, dst, 16, w, h, mode == 47);
	}
	printf("]}\n");
	return 0;
}

int main(int argc, char **argv) {
	if (argc > 1 && !strcmp(argv[1], "fuzz")) return fuzz_inter() != 0;
	if (argc > 1 && !strcmp(argv[1], "dump")) return dump_kat();
	fprintf(stderr, "usage: ref_kat fuzz\n");
	return 2;
}
