/* e264_decode — decode an Annex-B file through the edge264 C API and dump / hash / time it.
 * The SAME source is linked against the reference library (-> oracle/_ref/ref_decode) and against
 * libedge264_b200.so (-> tools/b200_decode): the loop below is the application-side contract
 * (reference README.md:117-156).  Test/bench infrastructure, not product.
 *   e264_decode in.264 [-o out.yuv] [-c] [-b reps] [-t n_threads] [-q]
 *     -o  write cropped planar I420 of every output frame      -c  print one FNV-1a hash per frame
 *     -b  decode the file `reps` times, print best & median wall seconds of the decode loop
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include "edge264.h"

static uint64_t fnv(uint64_t h, const uint8_t *p, int n) {
	for (int i = 0; i < n; i++) h = (h ^ p[i]) * 0x100000001b3ull;
	return h;
}
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static int cmpd(const void *a, const void *b) { double x = *(double*)a, y = *(double*)b; return (x > y) - (x < y); }

static int decode_once(const uint8_t *buf, size_t size, int n_threads, FILE *out, int hash, int quiet, int *last_ret, uint64_t *sig) {
	const uint8_t *end = buf + size;
	const uint8_t *nal = buf + 3 + (buf[2] == 0);
	Edge264Decoder *dec = edge264_alloc(n_threads, NULL, NULL, 0, NULL, NULL, NULL);
	if (!dec) { fprintf(stderr, "edge264_alloc failed\n"); exit(2); }
	Edge264Frame f;
	int res, frames = 0, drained = 0;
	uint64_t all = 0xcbf29ce484222325ull;
	for (;;) {
		const uint8_t *sc = nal < end ? edge264_find_start_code(nal, end, 0) : end;
		int frames_before = frames;
		res = edge264_decode_NAL(dec, nal, sc, NULL, NULL);
		if (nal >= end) drained = 1;
		while (!edge264_get_frame(dec, &f, 0)) {
			uint64_t h = 0xcbf29ce484222325ull;
			for (int y = 0; y < f.height_Y; y++) {
				const uint8_t *r = f.samples[0] + (size_t)y * f.stride_Y;
				if (out) fwrite(r, 1, f.width_Y, out);
				if (hash || sig) h = fnv(h, r, f.width_Y);
			}
			for (int c = 1; c < 3; c++) for (int y = 0; y < f.height_C; y++) {
				const uint8_t *r = f.samples[c] + (size_t)y * f.stride_C;
				if (out) fwrite(r, 1, f.width_C, out);
				if (hash || sig) h = fnv(h, r, f.width_C);
			}
			if (hash && !quiet) printf("frame %d id %d %dx%d hash %016llx\n", frames, f.FrameId, f.width_Y, f.height_Y, (unsigned long long)h);
			all = fnv(all, (uint8_t*)&h, 8);
			frames++;
		}
		if (res == ENOBUFS) {
			if (frames == frames_before) { if (!quiet) fprintf(stderr, "ENOBUFS with no frame to get: giving up\n"); break; }
			continue;
		}
		if (res != 0 && res != ENOTSUP && res != EBADMSG && !drained && !quiet)
			fprintf(stderr, "decode_NAL -> %d at offset %zu\n", res, (size_t)(nal - buf));
		if ((res == ENOTSUP || res == EBADMSG) && !quiet)
			fprintf(stderr, "decode_NAL -> %s at offset %zu\n", res == ENOTSUP ? "ENOTSUP" : "EBADMSG", (size_t)(nal - buf));
		if (res == ENOTSUP || res == EBADMSG) res = 0;
		nal = sc + 3 < end ? sc + 3 : end;
		if (sc < end && nal == end) nal = end;
		if (res != 0 || drained) break;
	}
	*last_ret = res;
	if (sig) *sig = all;
	edge264_free(&dec);
	return frames;
}

int main(int argc, char **argv) {
	const char *in = NULL, *outp = NULL; int hash = 0, reps = 0, nthreads = 0, quiet = 0;
	for (int i = 1; i < argc; i++) {
		if (!strcmp(argv[i], "-o")) outp = argv[++i];
		else if (!strcmp(argv[i], "-c")) hash = 1;
		else if (!strcmp(argv[i], "-q")) quiet = 1;
		else if (!strcmp(argv[i], "-b")) reps = atoi(argv[++i]);
		else if (!strcmp(argv[i], "-t")) nthreads = atoi(argv[++i]);
		else in = argv[i];
	}
	if (!in) { fprintf(stderr, "usage: %s in.264 [-o out.yuv] [-c] [-b reps] [-t threads]\n", argv[0]); return 2; }
	int fd = open(in, O_RDONLY); struct stat st;
	if (fd < 0 || fstat(fd, &st)) { perror(in); return 2; }
	uint8_t *buf = mmap(NULL, st.st_size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
	int ret; uint64_t sig;
	if (reps > 0) {
		double t[64]; int frames = 0; if (reps > 64) reps = 64;
		for (int r = 0; r < reps; r++) { double t0 = now(); frames = decode_once(buf, st.st_size, nthreads, NULL, 0, 1, &ret, NULL); t[r] = now() - t0; }
		qsort(t, reps, sizeof(double), cmpd);
		printf("{\"frames\": %d, \"best_s\": %.6f, \"median_s\": %.6f, \"fps_median\": %.3f}\n", frames, t[0], t[reps/2], frames / t[reps/2]);
		return 0;
	}
	FILE *out = outp ? fopen(outp, "wb") : NULL;
	int frames = decode_once(buf, st.st_size, nthreads, out, hash, quiet, &ret, &sig);
	if (out) fclose(out);
	printf("frames %d last_ret %d sig %016llx\n", frames, ret, (unsigned long long)sig);
	return 0;
}
