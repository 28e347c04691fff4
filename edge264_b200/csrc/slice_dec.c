/* slice_dec.c — the product-side instantiation of the shared slice-data code (parser direction). */
#include "mb_impl.h"

int e264_parse_slice_data(SliceCtx *s) {
	static int ready;
	if (!ready) { sx_init_tables(); ready = 1; }
	s->skip_run = -1; s->last_qp_delta_nz = 0; s->prev_mb_skipped = 0;
	if (s->cabac) {
		size_t byte = (s->br.pos + 7) >> 3;                 /* cabac_alignment_one_bit */
		if (byte > s->br.size) { s->error = 1; return 0; }
		cabac_init_states(s->cd.state, s->cabac_init_idc_col, s->qp);
		cabac_dec_start(&s->cd, s->br.buf + byte, s->br.buf + s->br.size);
	}
	int total = s->w_mbs * s->h_mbs, n = 0;
	for (;;) {
		if (s->mbaddr >= total) { s->error = 1; break; }
		int end = sx_one_mb(s);
		if (s->error) break;
		n++;
		if (s->cabac && cabac_dec_overrun(&s->cd)) { s->error = 1; break; }
		if (!s->cabac && s->br.overrun) { s->error = 1; break; }
		if (end) break;
		s->mbaddr++;
	}
	return n;
}
