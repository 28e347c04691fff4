/* records.h — the host→device record ABI of the reconstruction path.
 *
 * The CPU parser turns each macroblock into ONE fixed-size E264MbRec plus a variable run of int16
 * coefficients in the picture's coefficient pool; a picture also carries a small table of
 * E264SliceRec.  This replaces the reference's in-memory contract between its slice parser and its
 * pixel functions (reference: Edge264Macroblock edge264_internal.h:128-143, Edge264Task :223-261,
 * ctx->c[] :311) — the information is the same, the layout is ours.
 *
 * Index conventions: every per-4x4 quantity (mv, coded bits, intra modes) is indexed by the spec's
 * luma4x4BlkIdx "z-order" b = 4*(8x8 index) + (4x4 index inside it); 8x8 quantities by raster
 * 8x8 index.  Coefficients inside a block are in RASTER order (c[y*N+x]) as int16 levels before
 * dequantisation.
 *
 * Coefficient run of one macroblock, at pool[coef_off], in this order (all pieces are multiples of
 * 8 int16 so every block stays 16-byte aligned):
 *   I_PCM:  384 bytes of raw samples (Y 256, Cb 64, Cr 64) = 192 int16 slots
 *   else:   [16 luma DC levels, raster 4x4 over the 16 blocks]     if coded bit 24 (Intra16x16 only)
 *           for each coded luma 4x4 block in z-order: 16 levels    (transform 4x4)
 *           or for each coded 8x8 block: 64 levels                 (transform 8x8; its 4 bits are set together)
 *           [4 Cb DC + 4 Cr DC levels, raster 2x2]                 if coded bit 25 or 26
 *           for each coded chroma AC block (Cb0..3 then Cr0..3): 16 levels (slot 0 unused, zero)
 */
#ifndef E264B_RECORDS_H
#define E264B_RECORDS_H
#include <stdint.h>

enum { MBK_I4x4 = 0, MBK_I8x8 = 1, MBK_I16x16 = 2, MBK_IPCM = 3, MBK_INTER = 4 };

/* E264MbRec.flags */
#define MBF_T8x8     0x01   /* transform_size_8x8_flag */
#define MBF_EDGE_L   0x02   /* deblock the left macroblock edge  (reference filter_edges bit0, slice.c:1692,1724) */
#define MBF_EDGE_T   0x04   /* deblock the top macroblock edge   (bit1, slice.c:1762) */
#define MBF_DEBLOCK  0x08   /* this macroblock is deblocked at all (disable_deblocking_filter_idc != 1) */
#define MBF_SKIP     0x10   /* P_Skip / B_Skip (informational) */

/* bits of E264MbRec.coded */
#define CODED_LUMA(b)   (1u << (b))          /* b = luma4x4BlkIdx 0..15 */
#define CODED_CB_AC(i)  (1u << (16 + (i)))
#define CODED_CR_AC(i)  (1u << (20 + (i)))
#define CODED_Y_DC      (1u << 24)
#define CODED_CB_DC     (1u << 25)
#define CODED_CR_DC     (1u << 26)

/* intra mode bytes: low nibble = prediction mode of the standard (Intra4x4/8x8: 0..8, Intra16x16
 * and chroma: 0..3), high nibble = unavailability of the neighbours A(left)=1, B(top)=2,
 * C(top-right)=4, D(top-left)=8 — the device applies the substitution rules of 8.3.1.2 / 8.3.2.2
 * (the reference resolves the same information into its own 14/32/7/7 mode enums on the host:
 * slice.c:573-594, 720-725, 874-879). */
#define IMODE(mode, unavail) ((uint8_t)((mode) | ((unavail) << 4)))

typedef struct E264MbRec {
	uint8_t  kind;          /* MBK_* */
	uint8_t  flags;         /* MBF_* */
	uint8_t  qp[3];         /* QP_Y, QP'c for Cb, Cr of THIS macroblock */
	uint8_t  chroma_mode;   /* IMODE(intra_chroma_pred_mode, unavail A|B|D) */
	uint8_t  i16_mode;      /* IMODE(Intra16x16PredMode, unavail A|B|D) */
	uint8_t  slice_idx;     /* index into the picture's E264SliceRec table */
	uint32_t coded;         /* CODED_* */
	uint32_t coef_off;      /* int16 index into the picture's coefficient pool */
	uint8_t  modes[16];     /* IMODE per luma4x4BlkIdx (Intra4x4) or in [0..3] per 8x8 (Intra8x8) */
	int8_t   ref_idx[2][4]; /* per list, per 8x8: reference index, -1 = list unused */
	int8_t   ref_pic[2][4]; /* per list, per 8x8: frame slot of the reference picture, -1 = unused */
	uint8_t  reserved[16];
	int16_t  mv[2][16][2];  /* per list, per luma4x4BlkIdx: quarter-pel (x, y) */
} E264MbRec;                /* 192 bytes */

enum { WP_DEFAULT = 0, WP_EXPLICIT = 1, WP_IMPLICIT = 2 };

typedef struct E264SliceRec {
	int8_t   filter_offset_a;   /* FilterOffsetA = slice_alpha_c0_offset_div2 * 2 */
	int8_t   filter_offset_b;
	uint8_t  deblock_idc;       /* disable_deblocking_filter_idc */
	uint8_t  slice_type;        /* 0 P, 1 B, 2 I */
	uint8_t  wp_mode;           /* WP_* actually in force for this slice */
	uint8_t  luma_log2_wd;
	uint8_t  chroma_log2_wd;
	uint8_t  reserved;
	int16_t  wp_w[2][16][3];    /* explicit weights  [list][refIdx][Y,Cb,Cr] */
	int16_t  wp_o[2][16][3];    /* explicit offsets */
	int16_t  implicit_w1[16][16]; /* implicit bi-pred weight of the list-1 sample for (refIdxL0, refIdxL1), -64..128; w0 = 64 - w1 */
	uint8_t  scaling4x4[6][16]; /* merged scaling lists Y-intra,Cb-intra,Cr-intra,Y-inter,Cb-inter,Cr-inter; raster */
	uint8_t  scaling8x8[2][64]; /* Y-intra, Y-inter; raster */
} E264SliceRec;                 /* 8 + 192 + 192 + 512 + 96 + 128 = 1128 bytes */

/* One picture worth of work for the device (host view; device pointers are filled by the runtime). */
typedef struct E264PicDesc {
	int32_t width_mbs, height_mbs;
	int32_t stride_y, stride_c;   /* bytes; chroma rows alternate Cb,Cr inside one stride_c (reference headers.c:2027-2046) */
	int32_t plane_y;              /* byte offset of the chroma plane inside a frame buffer */
	int32_t frame_bytes;          /* size of one frame slot */
	int32_t dst_slot;             /* frame slot being reconstructed */
	int32_t n_slices;
	int32_t n_coefs;              /* int16 entries used in the coefficient pool */
	int32_t any_deblock;          /* at least one macroblock has MBF_DEBLOCK */
	int32_t n_intra;              /* number of intra macroblocks = entries of the staging's intra_list */
	int32_t staging;              /* handle of the staging area the picture was written into (E264Staging.handle) */
} E264PicDesc;

/* Host-visible staging of one picture, handed out by the backend (pinned memory for the CUDA backend): the parser
 * writes records, levels and slice records straight into it; intra_list receives the addresses of the picture's intra
 * macroblocks in raster order (the device schedules them from it). */
typedef struct E264Staging {
	int32_t handle;
	uint32_t coef_capacity;       /* int16 entries in coefs */
	struct E264MbRec *recs;
	int16_t *coefs;
	struct E264SliceRec *slices;
	uint32_t *intra_list;
} E264Staging;

static inline int e264_blk_x(int b) { return ((b & 1) | ((b >> 1) & 2)); }          /* z-order -> 4x4 column */
static inline int e264_blk_y(int b) { return (((b >> 1) & 1) | ((b >> 2) & 2)); }   /* z-order -> 4x4 row    */
static inline int e264_blk_z(int x, int y) { return (x & 1) | ((y & 1) << 1) | ((x & 2) << 1) | ((y & 2) << 2); }

#endif
