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
    if (rc == ORC_JPEG_OK) {
        if (J.nc == 1) {
            for (int y = 0; y < J.H; y++) memcpy(out + (size_t)y * J.W, J.c[0].plane + (size_t)y * J.c[0].pw, J.W);
        } else {
            for (int y = 0; y < J.H; y++)
                for (int x = 0; x < J.W; x++) {
                    const int Y = J.c[0].plane[(size_t)y * J.c[0].pw + x];
                    const int cb = chroma_at(&J, &J.c[1], x, y) - 128, cr = chroma_at(&J, &J.c[2], x, y) - 128;
                    uint8_t *o = out + ((size_t)y * J.W + x) * 3;
                    o[0] = sat8(Y + ((91881 * cr + 32768) >> 16));
                    o[1] = sat8(Y + ((-22554 * cb + 32768 - 46802 * cr) >> 16));
                    o[2] = sat8(Y + ((116130 * cb + 32768) >> 16));
                }
        }
    }
    for (int i = 0; i < J.nc; i++) free(J.c[i].plane);
    return rc;
}
