/*
 * oracle/jpeg_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
 *
 * Plain-C, strictly sequential restatement of what `np.array(Image.open(image))` computes for a baseline JPEG at
 * /root/reference/s1_lucaskanade_tracking.py:310 (s0_1_test_lucaskanade_tracking.py:79): SURVEY.md 8(a) row a1 / 8(f) rank 1.
 * The arithmetic lives in a third-party dependency that is not vendored in the reference: Pillow's bundled
 * libjpeg-turbo (this image: Pillow 12.2, libjpeg-turbo with the libjpeg v6.2 API), default decompression settings
 * (JDCT_ISLOW, fancy upsampling on, no smoothing).  Its published algorithm is restated here:
 *   marker parsing          ITU-T T.81 Annex B (SOF0/SOF1, DQT, DHT, SOS, DRI)
 *   Huffman decoding        T.81 Annex F.2.2 (canonical codes, EXTEND), zero runs / EOB / ZRL
 *   inverse DCT             libjpeg "islow" (Loeffler-Ligtenberg-Moshovitz, 13-bit constants, PASS1_BITS 2)
 *   chroma upsampling       libjpeg "fancy" triangle filter (h2v1: 3/4+1/4; h2v2: 9/16+3/16+3/16+1/16, biases 8/7),
 *                           edges replicate; plain replication when the chroma plane is <= 2 samples wide
 *   YCbCr -> RGB            libjpeg 16-bit fixed point (1.40200, 1.77200, 0.71414, 0.34414)
 * Pinning: tests/test_jpeg_oracle.py compares this decoder byte for byte with Pillow on the committed JPEG
 * fixtures (tests/golden/jpeg/, made by tests/golden/make_jpeg_golden.py) and on files encoded at test time.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_JPEG_OK 0
#define ORC_JPEG_E_FORMAT (-1)       /* not a JPEG / truncated header */
#define ORC_JPEG_E_UNSUPPORTED (-5)  /* progressive, arithmetic, 12-bit, CMYK, exotic sampling */

static const uint8_t ZIGZAG[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                   41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                   30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

typedef struct {
    int present;
    uint8_t bits[17], vals[256];
    int mincode[17], maxcode[17], valptr[17];
} htab_t;

typedef struct {
    int id, h, v, tq, td, ta;
    int pw, ph;          /* plane size in samples (whole blocks) */
    int dw, dh;          /* real (downsampled) size */
    uint8_t *plane;
    int pred;
} comp_t;

typedef struct {
    int W, H, nc;
    comp_t c[3];
    uint16_t q[4][64];   /* natural order */
    int qpresent[4];
    htab_t dc[4], ac[4];
    int ri;
    int hmax, vmax, mcux, mcuy;
    const uint8_t *scan;
    long scan_len;
} jpg_t;

static void build_htab(htab_t *t)
{
    int code = 0, k = 0;
    for (int l = 1; l <= 16; l++) {
        t->mincode[l] = code;
        t->valptr[l] = k;
        code += t->bits[l];
        k += t->bits[l];
        t->maxcode[l] = t->bits[l] ? code - 1 : -1;
        code <<= 1;
    }
    t->present = 1;
}

static int parse(const uint8_t *d, long n, jpg_t *J)
{
    memset(J, 0, sizeof(*J));
    if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return ORC_JPEG_E_FORMAT;
    long p = 2;
    int have_sof = 0;
    for (;;) {
        if (p + 4 > n) return ORC_JPEG_E_FORMAT;
        if (d[p] != 0xFF) return ORC_JPEG_E_FORMAT;
        while (p < n && d[p] == 0xFF) p++;                 /* fill bytes */
        if (p >= n) return ORC_JPEG_E_FORMAT;
        const int m = d[p++];
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
        if (m == 0xD9) return ORC_JPEG_E_FORMAT;
        if (p + 2 > n) return ORC_JPEG_E_FORMAT;
        const int len = (d[p] << 8) | d[p + 1];
        if (len < 2 || p + len > n) return ORC_JPEG_E_FORMAT;
        const uint8_t *s = d + p + 2;
        const int sl = len - 2;
        if (m == 0xC0 || m == 0xC1) {
            if (sl < 6) return ORC_JPEG_E_FORMAT;
            if (s[0] != 8) return ORC_JPEG_E_UNSUPPORTED;
            J->H = (s[1] << 8) | s[2];
            J->W = (s[3] << 8) | s[4];
            J->nc = s[5];
            if (J->nc != 1 && J->nc != 3) return ORC_JPEG_E_UNSUPPORTED;
            if (sl < 6 + 3 * J->nc || J->W == 0 || J->H == 0) return ORC_JPEG_E_FORMAT;
            for (int i = 0; i < J->nc; i++) {
                J->c[i].id = s[6 + 3 * i];
                J->c[i].h = s[7 + 3 * i] >> 4;
                J->c[i].v = s[7 + 3 * i] & 15;
                J->c[i].tq = s[8 + 3 * i] & 3;
            }
            have_sof = 1;
        } else if (m == 0xC2 || m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC)) {
            return ORC_JPEG_E_UNSUPPORTED;
        } else if (m == 0xC4) {
            int o = 0;
            while (o < sl) {
                if (o + 17 > sl) return ORC_JPEG_E_FORMAT;
                const int tc = s[o] >> 4, th = s[o] & 15;
                if (tc > 1 || th > 3) return ORC_JPEG_E_FORMAT;
                htab_t *t = tc ? &J->ac[th] : &J->dc[th];
                memset(t, 0, sizeof(*t));
                int cnt = 0;
                for (int l = 1; l <= 16; l++) { t->bits[l] = s[o + l]; cnt += s[o + l]; }
                o += 17;
                if (cnt > 256 || o + cnt > sl) return ORC_JPEG_E_FORMAT;
                memcpy(t->vals, s + o, cnt);
                o += cnt;
                build_htab(t);
            }
        } else if (m == 0xDB) {
            int o = 0;
            while (o < sl) {
                const int pq = s[o] >> 4, tq = s[o] & 15;
                if (tq > 3 || pq > 1) return ORC_JPEG_E_FORMAT;
                o++;
                if (o + 64 * (pq + 1) > sl) return ORC_JPEG_E_FORMAT;
                for (int i = 0; i < 64; i++) {
                    const int v = pq ? ((s[o] << 8) | s[o + 1]) : s[o];
                    o += pq + 1;
                    J->q[tq][ZIGZAG[i]] = (uint16_t)v;
                }
                J->qpresent[tq] = 1;
            }
        } else if (m == 0xDD) {
            if (sl < 2) return ORC_JPEG_E_FORMAT;
            J->ri = (s[0] << 8) | s[1];
        } else if (m == 0xDA) {
            if (!have_sof || sl < 1) return ORC_JPEG_E_FORMAT;
            const int ns = s[0];
            if (ns != J->nc || sl < 1 + 2 * ns + 3) return ORC_JPEG_E_UNSUPPORTED;
            for (int i = 0; i < ns; i++) {
                if (s[1 + 2 * i] != J->c[i].id) return ORC_JPEG_E_UNSUPPORTED;
                J->c[i].td = s[2 + 2 * i] >> 4;
                J->c[i].ta = s[2 + 2 * i] & 15;
                if (J->c[i].td > 3 || J->c[i].ta > 3) return ORC_JPEG_E_FORMAT;
            }
            J->scan = d + p + len;
            J->scan_len = n - (p + len);
            break;
        }
        p += len;
    }
    if (J->nc == 1) { J->c[0].h = 1; J->c[0].v = 1; }
    J->hmax = J->vmax = 1;
    for (int i = 0; i < J->nc; i++) {
        if (J->c[i].h < 1 || J->c[i].v < 1) return ORC_JPEG_E_FORMAT;
        if (J->c[i].h > J->hmax) J->hmax = J->c[i].h;
        if (J->c[i].v > J->vmax) J->vmax = J->c[i].v;
        if (!J->qpresent[J->c[i].tq] || !J->dc[J->c[i].td].present || !J->ac[J->c[i].ta].present) return ORC_JPEG_E_FORMAT;
    }
    if (J->nc == 3) {
        /* supported: chroma 1x1 under luma 1x1 (4:4:4), 2x1 (4:2:2), 2x2 (4:2:0) */
        if (J->c[1].h != 1 || J->c[1].v != 1 || J->c[2].h != 1 || J->c[2].v != 1) return ORC_JPEG_E_UNSUPPORTED;
        if (!((J->hmax == 1 && J->vmax == 1) || (J->hmax == 2 && J->vmax == 1) || (J->hmax == 2 && J->vmax == 2)))
            return ORC_JPEG_E_UNSUPPORTED;
    }
    J->mcux = (J->W + 8 * J->hmax - 1) / (8 * J->hmax);
    J->mcuy = (J->H + 8 * J->vmax - 1) / (8 * J->vmax);
    for (int i = 0; i < J->nc; i++) {
        comp_t *c = &J->c[i];
        c->pw = J->mcux * c->h * 8;
        c->ph = J->mcuy * c->v * 8;
        c->dw = (J->W * c->h + J->hmax - 1) / J->hmax;
        c->dh = (J->H * c->v + J->vmax - 1) / J->vmax;
    }
    return ORC_JPEG_OK;
}

/* ---- bit reader (T.81 F.2.2.5, byte stuffing B.1.1.5) ---- */
typedef struct {
    const uint8_t *d;
    long n, pos;
    uint32_t acc;
    int nbits;
    int marker;          /* a marker was hit: feed zero bits */
} bits_t;

static void fill(bits_t *b)
{
    while (b->nbits <= 24) {
        int byte = 0;
        if (!b->marker && b->pos < b->n) {
            byte = b->d[b->pos];
            if (byte == 0xFF) {
                if (b->pos + 1 < b->n && b->d[b->pos + 1] == 0x00) b->pos += 2;
                else { b->marker = 1; byte = 0; }
            } else b->pos++;
        }
        b->acc |= (uint32_t)byte << (24 - b->nbits);
        b->nbits += 8;
    }
}
static int getbits(bits_t *b, int k)
{
    if (k == 0) return 0;
    fill(b);
    const int v = (int)(b->acc >> (32 - k));
    b->acc <<= k;
    b->nbits -= k;
    return v;
}
static int decode_sym(bits_t *b, const htab_t *t)
{
    int code = 0;
    for (int l = 1; l <= 16; l++) {
        code = (code << 1) | getbits(b, 1);
        if (t->maxcode[l] >= 0 && code <= t->maxcode[l] && code >= t->mincode[l]) return t->vals[t->valptr[l] + code - t->mincode[l]];
    }
    return 0;
}
static int extend(int r, int s) { return r < (1 << (s - 1)) ? r - (1 << s) + 1 : r; }

/* ---- libjpeg jidctint.c (islow), dequantisation folded in ---- */
#define CONST_BITS 13
#define PASS1_BITS 2
#define DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))
static void idct_1d(const int *in, int stride, int *o, int pass)
{
    int z2 = in[2 * stride], z3 = in[6 * stride];
    int z1 = (z2 + z3) * 4433;
    int tmp2 = z1 + z3 * (-15137);
    int tmp3 = z1 + z2 * 6270;
    z2 = in[0];
    z3 = in[4 * stride];
    int tmp0 = (int)((unsigned)(z2 + z3) << CONST_BITS);
    int tmp1 = (int)((unsigned)(z2 - z3) << CONST_BITS);
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = in[7 * stride]; tmp1 = in[5 * stride]; tmp2 = in[3 * stride]; tmp3 = in[1 * stride];
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int z4 = tmp1 + tmp3;
    const int z5 = (z3 + z4) * 9633;
    tmp0 *= 2446; tmp1 *= 16819; tmp2 *= 25172; tmp3 *= 12299;
    z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    const int sh = pass == 1 ? CONST_BITS - PASS1_BITS : CONST_BITS + PASS1_BITS + 3;
    o[0] = DESCALE(tmp10 + tmp3, sh); o[7] = DESCALE(tmp10 - tmp3, sh);
    o[1] = DESCALE(tmp11 + tmp2, sh); o[6] = DESCALE(tmp11 - tmp2, sh);
    o[2] = DESCALE(tmp12 + tmp1, sh); o[5] = DESCALE(tmp12 - tmp1, sh);
    o[3] = DESCALE(tmp13 + tmp0, sh); o[4] = DESCALE(tmp13 - tmp0, sh);
}
static void idct_block(const int16_t *coef, const uint16_t *q, uint8_t *dst, int pitch)
{
    int in[64], ws[64], o[8];
    for (int i = 0; i < 64; i++) in[i] = (int)coef[i] * (int)q[i];
    for (int c = 0; c < 8; c++) {
        idct_1d(in + c, 8, o, 1);
        for (int r = 0; r < 8; r++) ws[r * 8 + c] = o[r];
    }
    for (int r = 0; r < 8; r++) {
        idct_1d(ws + r * 8, 1, o, 2);
        for (int c = 0; c < 8; c++) {
            int v = o[c] + 128;
            dst[r * pitch + c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
}

static int decode_scan(jpg_t *J)
{
    bits_t b = {J->scan, J->scan_len, 0, 0, 0, 0};
    int16_t coef[64];
    int since_restart = 0;
    for (int i = 0; i < J->nc; i++) J->c[i].pred = 0;
    for (int my = 0; my < J->mcuy; my++)
        for (int mx = 0; mx < J->mcux; mx++) {
            if (J->ri && since_restart == J->ri) {
                /* RSTn: drop the padding bits, skip the marker, reset the predictions */
                b.acc = 0; b.nbits = 0;
                if (b.marker) { b.marker = 0; }
                while (b.pos + 1 < b.n && !(b.d[b.pos] == 0xFF && b.d[b.pos + 1] >= 0xD0 && b.d[b.pos + 1] <= 0xD7)) b.pos++;
                b.pos += 2;
                for (int i = 0; i < J->nc; i++) J->c[i].pred = 0;
                since_restart = 0;
            }
            since_restart++;
            for (int ci = 0; ci < J->nc; ci++) {
                comp_t *c = &J->c[ci];
                for (int v = 0; v < c->v; v++)
                    for (int h = 0; h < c->h; h++) {
                        memset(coef, 0, sizeof(coef));
                        int s = decode_sym(&b, &J->dc[c->td]);
                        int diff = 0;
                        if (s) diff = extend(getbits(&b, s), s);
                        c->pred += diff;
                        coef[0] = (int16_t)c->pred;
                        for (int k = 1; k < 64; k++) {
                            const int rs = decode_sym(&b, &J->ac[c->ta]);
                            const int r = rs >> 4;
                            s = rs & 15;
                            if (s) {
                                k += r;
                                const int val = extend(getbits(&b, s), s);
                                if (k < 64) coef[ZIGZAG[k]] = (int16_t)val;
                            } else {
                                if (r != 15) break;
                                k += 15;
                            }
                        }
                        idct_block(coef, J->q[c->tq], c->plane + (size_t)((my * c->v + v) * 8) * c->pw + (mx * c->h + h) * 8, c->pw);
                    }
            }
        }
    return ORC_JPEG_OK;
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static uint8_t sat8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

/* chroma sample at full-resolution (x, y): libjpeg jdsample.c */
static int chroma_at(const jpg_t *J, const comp_t *c, int x, int y)
{
    const int hs = J->hmax / c->h, vs = J->vmax / c->v;
    const uint8_t *P = c->plane;
    const int pw = c->pw;
    if (hs == 1 && vs == 1) return P[(size_t)y * pw + x];
    const int cx = x >> 1;
    if (hs == 2 && vs == 1) {
        if (c->dw <= 2) return P[(size_t)y * pw + cx];                            /* h2v1_upsample (replicate) */
        const int t = P[(size_t)y * pw + cx];
        if (x & 1) return (3 * t + P[(size_t)y * pw + clampi(cx + 1, 0, c->dw - 1)] + 2) >> 2;
        return (3 * t + P[(size_t)y * pw + clampi(cx - 1, 0, c->dw - 1)] + 1) >> 2;
    }
    const int cy = y >> 1;
    if (c->dw <= 2) return P[(size_t)cy * pw + cx];                               /* h2v2_upsample (replicate) */
    const int fy = clampi((y & 1) ? cy + 1 : cy - 1, 0, c->dh - 1);
    const uint8_t *r0 = P + (size_t)cy * pw, *r1 = P + (size_t)fy * pw;
    const int cs = 3 * r0[cx] + r1[cx];
    const int ox = clampi((x & 1) ? cx + 1 : cx - 1, 0, c->dw - 1);
    const int os = 3 * r0[ox] + r1[ox];
    return (3 * cs + os + ((x & 1) ? 7 : 8)) >> 4;
}

/* planes -> (H, W, 3) RGB u8 / (H, W) u8: fancy upsampling + jdcolor.c */
static void emit_pixels(const jpg_t *J, uint8_t *out)
{
    if (J->nc == 1) {
        for (int y = 0; y < J->H; y++) memcpy(out + (size_t)y * J->W, J->c[0].plane + (size_t)y * J->c[0].pw, J->W);
        return;
    }
    for (int y = 0; y < J->H; y++)
        for (int x = 0; x < J->W; x++) {
            const int Y = J->c[0].plane[(size_t)y * J->c[0].pw + x];
            const int cb = chroma_at(J, &J->c[1], x, y) - 128, cr = chroma_at(J, &J->c[2], x, y) - 128;
            uint8_t *o = out + ((size_t)y * J->W + x) * 3;
            o[0] = sat8(Y + ((91881 * cr + 32768) >> 16));
            o[1] = sat8(Y + ((-22554 * cb + 32768 - 46802 * cr) >> 16));
            o[2] = sat8(Y + ((116130 * cb + 32768) >> 16));
        }
}

int orc_jpeg_info(const uint8_t *file, long n, int *W, int *H, int *nc)
{
    jpg_t J;
    const int rc = parse(file, n, &J);
    if (rc) return rc;
    *W = J.W; *H = J.H; *nc = J.nc;
    return ORC_JPEG_OK;
}

/* out: (H, W, 3) RGB u8 for a colour file, (H, W) u8 for a grey-scale file -- what np.array(Image.open(f)) holds */
int orc_jpeg_decode(const uint8_t *file, long n, uint8_t *out)
{
    jpg_t J;
    int rc = parse(file, n, &J);
    if (rc) return rc;
    for (int i = 0; i < J.nc; i++) {
        J.c[i].plane = (uint8_t *)calloc((size_t)J.c[i].pw * J.c[i].ph, 1);
        if (!J.c[i].plane) return ORC_JPEG_E_FORMAT;
    }
    rc = decode_scan(&J);
    if (rc == ORC_JPEG_OK) emit_pixels(&J, out);
    for (int i = 0; i < J.nc; i++) free(J.c[i].plane);
    return rc;
}

/* ================================================================================================================
 * Save-and-reopen round trip of the reference's cropping pre-pass: `img_crop.save(outpath)` (imports/camtools.py:80,102,
 * 232) followed by `np.array(Image.open(image))` (s1_lucaskanade_tracking.py:310).  Pillow's save with default settings
 * = libjpeg-turbo compression at quality 75, 4:2:0, JDCT_ISLOW, no smoothing; entropy coding is lossless, so the pixels
 * the tracker sees are: decode( quantise( FDCT( downsample( RGB->YCbCr(crop) )))).  Restated from the published libjpeg
 * algorithm (jccolor.c rgb_ycc_convert, jcprepct.c / jcsample.c edge replication + h2v2 / h2v1 box filters with the
 * alternating bias, jfdctint.c, jcdctmgr.c quantisation with divisor 8 * q and round-half-up on the magnitude,
 * jcparam.c quality scaling of the Annex K tables) + the decode half above.
 * Pinning: tests/test_jpeg_oracle.py compares this with Pillow's own save + open on random crops.
 * ================================================================================================================ */
static const uint8_t STD_LUMA_Q[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                                       14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                                       18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                                       49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const uint8_t STD_CHROMA_Q[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                                         24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                                         99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                         99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

/* jcparam.c: jpeg_quality_scaling + jpeg_add_quant_table(force_baseline = TRUE); tables in natural order */
void orc_jpeg_quality_tables(int quality, uint16_t *luma, uint16_t *chroma)
{
    if (quality <= 0) quality = 1;
    if (quality > 100) quality = 100;
    const int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
    for (int i = 0; i < 64; i++) {
        long a = ((long)STD_LUMA_Q[i] * scale + 50) / 100, b = ((long)STD_CHROMA_Q[i] * scale + 50) / 100;
        luma[i] = (uint16_t)(a < 1 ? 1 : (a > 255 ? 255 : a));
        chroma[i] = (uint16_t)(b < 1 ? 1 : (b > 255 ? 255 : b));
    }
}

/* jfdctint.c: one 8-point pass; pass 1 = rows (results scaled up by 2^PASS1_BITS), pass 2 = columns */
static void fdct_1d(int *d, int stride, int pass)
{
    const int t0 = d[0] + d[7 * stride], t7 = d[0] - d[7 * stride], t1 = d[stride] + d[6 * stride], t6 = d[stride] - d[6 * stride];
    const int t2 = d[2 * stride] + d[5 * stride], t5 = d[2 * stride] - d[5 * stride], t3 = d[3 * stride] + d[4 * stride],
              t4 = d[3 * stride] - d[4 * stride];
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    const int sh = pass == 1 ? CONST_BITS - PASS1_BITS : CONST_BITS + PASS1_BITS;
    if (pass == 1) {
        d[0] = (int)((unsigned)(t10 + t11) << PASS1_BITS);
        d[4 * stride] = (int)((unsigned)(t10 - t11) << PASS1_BITS);
    } else {
        d[0] = DESCALE(t10 + t11, PASS1_BITS);
        d[4 * stride] = DESCALE(t10 - t11, PASS1_BITS);
    }
    int z1 = (t12 + t13) * 4433;
    d[2 * stride] = DESCALE(z1 + t13 * 6270, sh);
    d[6 * stride] = DESCALE(z1 + t12 * (-15137), sh);
    z1 = t4 + t7;
    int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
    const int z5 = (z3 + z4) * 9633;
    const int a4 = t4 * 2446, a5 = t5 * 16819, a6 = t6 * 25172, a7 = t7 * 12299;
    z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
    z3 += z5; z4 += z5;
    d[7 * stride] = DESCALE(a4 + z1 + z3, sh);
    d[5 * stride] = DESCALE(a5 + z2 + z4, sh);
    d[3 * stride] = DESCALE(a6 + z2 + z3, sh);
    d[stride] = DESCALE(a7 + z1 + z4, sh);
}

/* one block in place: samples -> FDCT -> quantise (jcdctmgr.c) -> [entropy coding is lossless] -> dequantise -> IDCT */
static void requantise_block(uint8_t *blk, int pitch, const uint16_t *q)
{
    int w[64];
    int16_t coef[64];
    for (int r = 0; r < 8; r++)
        for (int c = 0; c < 8; c++) w[r * 8 + c] = (int)blk[r * pitch + c] - 128;
    for (int r = 0; r < 8; r++) fdct_1d(w + r * 8, 1, 1);
    for (int c = 0; c < 8; c++) fdct_1d(w + c, 8, 2);
    for (int i = 0; i < 64; i++) {
        const int qv = (int)q[i] << 3;
        int t = w[i];
        if (t < 0) { t = -t; t += qv >> 1; t = t >= qv ? t / qv : 0; t = -t; }
        else { t += qv >> 1; t = t >= qv ? t / qv : 0; }
        coef[i] = (int16_t)t;
    }
    idct_block(coef, q, blk, pitch);
}

/* rgb: (H, W, 3) u8 with row pitch `pitch` bytes; hs, vs = luma sampling factors (2,2 = Pillow's default 4:2:0; 2,1; 1,1);
 * out: (H, W, 3) u8, what np.array(Image.open(saved_file)) returns */
int orc_jpeg_recompress(const uint8_t *rgb, long pitch, int W, int H, int quality, int hs, int vs, uint8_t *out)
{
    if (W <= 0 || H <= 0 || W > 65535 || H > 65535) return ORC_JPEG_E_FORMAT;
    if (!((hs == 1 && vs == 1) || (hs == 2 && vs == 1) || (hs == 2 && vs == 2))) return ORC_JPEG_E_UNSUPPORTED;
    jpg_t J;
    memset(&J, 0, sizeof(J));
    J.W = W; J.H = H; J.nc = 3; J.hmax = hs; J.vmax = vs;
    J.c[0].h = hs; J.c[0].v = vs; J.c[1].h = J.c[1].v = J.c[2].h = J.c[2].v = 1;
    J.c[0].tq = 0; J.c[1].tq = J.c[2].tq = 1;
    orc_jpeg_quality_tables(quality, J.q[0], J.q[1]);
    J.mcux = (W + 8 * hs - 1) / (8 * hs);
    J.mcuy = (H + 8 * vs - 1) / (8 * vs);
    for (int i = 0; i < 3; i++) {
        comp_t *c = &J.c[i];
        c->pw = J.mcux * c->h * 8; c->ph = J.mcuy * c->v * 8;
        c->dw = (W * c->h + hs - 1) / hs; c->dh = (H * c->v + vs - 1) / vs;
        c->plane = (uint8_t *)calloc((size_t)c->pw * c->ph, 1);
        if (!c->plane) return ORC_JPEG_E_FORMAT;
    }
    /* jccolor.c rgb_ycc_convert at full resolution (16-bit fixed point) */
    uint8_t *ycc = (uint8_t *)malloc((size_t)W * H * 3);
    if (!ycc) return ORC_JPEG_E_FORMAT;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const uint8_t *p = rgb + (size_t)y * pitch + (size_t)x * 3;
            const int r = p[0], g = p[1], b = p[2];
            uint8_t *o = ycc + ((size_t)y * W + x) * 3;
            o[0] = (uint8_t)((19595 * r + 38470 * g + 7471 * b + 32768) >> 16);
            o[1] = (uint8_t)((-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16);
            o[2] = (uint8_t)((32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16);
        }
#define YCC(x, y, k) ycc[((size_t)((y) < H - 1 ? (y) : H - 1) * W + ((x) < W - 1 ? (x) : W - 1)) * 3 + (k)]
    /* luma: fullsize_downsample; rows and columns past the image replicate the last one (jcprepct.c, expand_right_edge) */
    for (int y = 0; y < J.c[0].ph; y++)
        for (int x = 0; x < J.c[0].pw; x++) J.c[0].plane[(size_t)y * J.c[0].pw + x] = YCC(x, y, 0);
    /* chroma: box filter on the edge-replicated full-resolution rows; rows of the plane past the last real row group copy it */
    for (int k = 1; k < 3; k++) {
        comp_t *c = &J.c[k];
        for (int cy = 0; cy < c->ph; cy++) {
            const int ce = cy < c->dh - 1 ? cy : c->dh - 1;
            for (int cx = 0; cx < c->pw; cx++) {
                int v;
                if (hs == 1) v = YCC(cx, ce, k);
                else if (vs == 1) v = (YCC(2 * cx, ce, k) + YCC(2 * cx + 1, ce, k) + (cx & 1)) >> 1;
                else v = (YCC(2 * cx, 2 * ce, k) + YCC(2 * cx + 1, 2 * ce, k) + YCC(2 * cx, 2 * ce + 1, k) + YCC(2 * cx + 1, 2 * ce + 1, k) + 1 + (cx & 1)) >> 2;
                c->plane[(size_t)cy * c->pw + cx] = (uint8_t)v;
            }
        }
    }
#undef YCC
    free(ycc);
    for (int i = 0; i < 3; i++) {
        comp_t *c = &J.c[i];
        for (int by = 0; by < c->ph / 8; by++)
            for (int bx = 0; bx < c->pw / 8; bx++) requantise_block(c->plane + (size_t)by * 8 * c->pw + bx * 8, c->pw, J.q[c->tq]);
    }
    emit_pixels(&J, out);
    for (int i = 0; i < 3; i++) free(J.c[i].plane);
    return ORC_JPEG_OK;
}
