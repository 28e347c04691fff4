/* oracle/ref_log.c — TEST INFRASTRUCTURE.  Decodes an Annex-B file with the UNMODIFIED reference decoder (logs variant,
 * macroblock logging on) and writes its YAML log: the input format of the reference's own stream generator
 * /root/reference/tests/gen_avc.py.  tests/make_gen_avc_fixture.py chains the two to obtain a stream whose every bit was
 * written by the reference's generator (BASELINE.json configs[0]).  Built into oracle/_ref/ by oracle/Makefile. */
#include <stdio.h>
#include <stdlib.h>
#include <errno.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include "edge264.h"
static int log_cb(const char *s, void *arg) { return fputs(s, (FILE *)arg); }
int main(int argc, char **argv) {
	if (argc < 3) { fprintf(stderr, "usage: %s in.264 out.yaml\n", argv[0]); return 2; }
	int fd = open(argv[1], O_RDONLY); struct stat st;
	if (fd < 0 || fstat(fd, &st)) { perror(argv[1]); return 2; }
	const uint8_t *buf = mmap(NULL, st.st_size, PROT_READ, MAP_PRIVATE, fd, 0), *end = buf + st.st_size;
	FILE *out = fopen(argv[2], "w");
	fputs("--- # log of the reference decoder\n", out);
	Edge264Decoder *dec = edge264_alloc(0, log_cb, out, 1, NULL, NULL, NULL);
	if (!dec) { fprintf(stderr, "edge264_alloc with logs failed\n"); return 2; }
	const uint8_t *nal = buf + 3 + (buf[2] == 0);
	Edge264Frame f; int res;
	do {
		const uint8_t *sc = edge264_find_start_code(nal, end, 0);
		res = edge264_decode_NAL(dec, nal, sc, NULL, NULL);
		while (!edge264_get_frame(dec, &f, 0));
		if (res == ENOBUFS) continue;
		if (res == ENOTSUP || res == EBADMSG) res = 0;
		nal = sc + 3;
	} while (res == 0 && nal < end);
	edge264_free(&dec);
	fclose(out);
	return 0;
}
