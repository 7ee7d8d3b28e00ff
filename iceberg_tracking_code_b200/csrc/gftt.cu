// K2: Shi-Tomasi corner seeding.  Replaces
// cv2.goodFeaturesToTrack(frame_gray, mask=mask, **feature_params), s1_lucaskanade_tracking.py:437
// (s0_1_test_lucaskanade_tracking.py:167).
//
//   K2a  eig_kernel      Sobel3 -> (sx^2, sx*sy, sy^2) as exact integers -> blockSize^2 box sum by a horizontal
//                        window sum + a vertical sliding sum (int32, exact, REFLECT_101 on the product image)
//                        -> lambda_min map + masked global max (ordered-uint atomicMax).  OpenCV rounds every
//                        product to float before its double box sum; the exact sums sit in the middle of that
//                        rounding noise (~1e-7 relative, same size as the wheel's own irreproducibility).
//   K2b  nms_kernel      threshold (q * max), 3x3 non-maximum suppression, mask, 1-px border; candidates appended
//                        as 64-bit keys (response bits << 32 | linear address)
//        radix sort      stable LSD, 8-bit digits, warp match_any ranking: all candidates ordered by
//                        (response desc, address desc) exactly like OpenCV's comparator -> rank
//   K2c  cells + rounds  OpenCV's sequential greedy min-distance culling restated as a fixed point that is safe to run
//                        in parallel: candidates are bucketed on OpenCV's cell grid (cell = round(minDistance)); a
//                        candidate is REJECTED once a stronger candidate within minDistance in its 3x3 cell
//                        neighbourhood is ACCEPTED, ACCEPTED once all such candidates are REJECTED.  Decisions are
//                        final and monotone, so rounds may read each other's fresh or stale states.
//        compaction      accepted candidates in rank order -> (x, y) float32, first maxCorners
#include "common.cuh"
#include <stddef.h>
#include <string.h>

namespace ibt {

// ---------------------------------------------------------------------------------------------
// K2a
constexpr int EW = 256;                 // threads per CTA = support columns per CTA row
constexpr int ERS = 64;                 // output rows per CTA
constexpr int MAX_BLOCK = 31;

__device__ __forceinline__ uint32_t enc_f32(float f)           // order-preserving float -> uint
{
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f32(uint32_t e)
{
    return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

__global__ void __launch_bounds__(EW)
eig_kernel(const uint8_t *__restrict__ img, int H, int W, int64_t pitch, int bs, double s2,
           float *__restrict__ eig, int64_t eig_pitch_f,
           const uint8_t *__restrict__ mask, int64_t mask_pitch, uint32_t *__restrict__ maxbits)
{
    extern __shared__ int eig_smem[];
    int *prod = eig_smem;                      // [2][3][EW]   products of the current support row (double-buffered)
    int *ring = eig_smem + 2 * 3 * EW;         // [bs][3][EW]  horizontal window sums of the last bs support rows
    const int t = threadIdx.x;
    const int nout = EW - bs + 1;              // output columns per CTA
    const int a0 = -(bs / 2);
    const int X0 = blockIdx.x * nout, Y0 = blockIdx.y * ERS;
    const int x = X0 + t;                      // output column of this thread (t < nout)
    const bool xout = t < nout && x < W;
    // REFLECT_101 acts on the product image: support column -> image column
    const int px = r101(X0 + a0 + t, W);
    const int xm = r101(px - 1, W), xp = r101(px + 1, W);
    int vs0 = 0, vs1 = 0, vs2 = 0;
    uint32_t lmax = 0;
    const int nrows = min(ERS, H - Y0) + bs - 1;
    int slot = 0;
    // the 3x3 bytes of support row k+1 are fetched before the barrier of row k: the loads fly during the window sums
    int b00, b01, b02, b10, b12, b20, b21, b22;
    {
        const int py = r101(Y0 + a0, H);
        const uint8_t *r0 = img + (int64_t)r101(py - 1, H) * pitch, *r1 = img + (int64_t)py * pitch;
        const uint8_t *r2 = img + (int64_t)r101(py + 1, H) * pitch;
        b00 = r0[xm]; b01 = r0[px]; b02 = r0[xp]; b10 = r1[xm]; b12 = r1[xp]; b20 = r2[xm]; b21 = r2[px]; b22 = r2[xp];
    }
    for (int k = 0; k < nrows; k++) {
        const int sx = (b02 + 2 * b12 + b22) - (b00 + 2 * b10 + b20);
        const int sy = (b20 + 2 * b21 + b22) - (b00 + 2 * b01 + b02);
        int *pb = prod + (k & 1) * 3 * EW;
        pb[t] = sx * sx; pb[EW + t] = sx * sy; pb[2 * EW + t] = sy * sy;
        if (k + 1 < nrows) {
            const int py = r101(Y0 + a0 + k + 1, H);
            const uint8_t *r0 = img + (int64_t)r101(py - 1, H) * pitch, *r1 = img + (int64_t)py * pitch;
            const uint8_t *r2 = img + (int64_t)r101(py + 1, H) * pitch;
            b00 = r0[xm]; b01 = r0[px]; b02 = r0[xp]; b10 = r1[xm]; b12 = r1[xp]; b20 = r2[xm]; b21 = r2[px]; b22 = r2[xp];
        }
        __syncthreads();
        if (t < nout) {
            int h0 = 0, h1 = 0, h2 = 0;
            for (int j = 0; j < bs; j++) { h0 += pb[t + j]; h1 += pb[EW + t + j]; h2 += pb[2 * EW + t + j]; }
            int *rg = ring + slot * 3 * EW;
            if (k >= bs) { vs0 -= rg[t]; vs1 -= rg[EW + t]; vs2 -= rg[2 * EW + t]; }
            vs0 += h0; vs1 += h1; vs2 += h2;
            rg[t] = h0; rg[EW + t] = h1; rg[2 * EW + t] = h2;
            if (k >= bs - 1 && xout) {
                const int y = Y0 + k - (bs - 1);
                const float a = __fmul_rn((float)((double)vs0 * s2), 0.5f), b = (float)((double)vs1 * s2);
                const float c = __fmul_rn((float)((double)vs2 * s2), 0.5f);
                const float d = __fsub_rn(a, c);
                const float e = __fsub_rn(__fadd_rn(a, c), __fsqrt_rn(__fadd_rn(__fmul_rn(d, d), __fmul_rn(b, b))));
                eig[(int64_t)y * eig_pitch_f + x] = e;
                if (maxbits && (!mask || mask[(int64_t)y * mask_pitch + x])) lmax = max(lmax, enc_f32(e));
            }
        }
        slot = slot + 1 == bs ? 0 : slot + 1;
    }
    if (maxbits) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        if ((t & 31) == 0 && lmax) atomicMax(maxbits, lmax);
    }
}

static size_t eig_smem_bytes(int bs) { return (size_t)(2 + bs) * 3 * EW * sizeof(int); }

static int launch_eig(const uint8_t *gray, int H, int W, int64_t pitch, int bs, float *eig, int64_t eig_pitch_bytes,
                      const uint8_t *mask, int64_t mask_pitch, uint32_t *maxbits, cudaStream_t st)
{
    if (!gray || !eig || H <= 0 || W <= 0 || bs < 1 || bs > MAX_BLOCK || pitch < W || eig_pitch_bytes % 4 != 0 ||
        eig_pitch_bytes < (int64_t)W * 4)
        return IBT_E_INVALID;
    if (eig_smem_bytes(bs) > 48 * 1024)                 // large blockSize: opt in to > 48 KB dynamic shared memory (per device)
        IBT_CUDA_TRY(cudaFuncSetAttribute(eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eig_smem_bytes(MAX_BLOCK)));
    const float scale = (float)(1.0 / (4.0 * bs * 255.0));       // OpenCV's Sobel scale for ksize 3
    const double s2 = (double)scale * (double)scale;
    const int nout = EW - bs + 1;
    dim3 grid((W + nout - 1) / nout, (H + ERS - 1) / ERS);
    eig_kernel<<<grid, EW, eig_smem_bytes(bs), st>>>(gray, H, W, pitch, bs, s2, eig, eig_pitch_bytes / 4, mask, mask_pitch, maxbits);
    return check_launch("eig_kernel");
}

// ---------------------------------------------------------------------------------------------
// K2b
enum : uint8_t { ST_UNDECIDED = 0, ST_ACCEPTED = 1, ST_REJECTED = 2 };

struct GfttCounters {
    uint32_t maxbits;        // enc_f32 of the masked maximum, 0 = no allowed pixel
    uint32_t ncand;          // candidates found (may exceed capacity)
    uint32_t nacc;           // accepted after culling
    uint32_t pad;
    uint32_t remaining[64];  // undecided candidates left after round k (mod 64)
    uint32_t hist[4096];     // candidates per top-12-bit bin of the ordered response (top-k prefilter)
};

__device__ __forceinline__ float tozero(float v, float thr) { return v > thr ? v : 0.f; }

// one thread per 4 x 4 pixel block: six row loads (float4 + the two neighbouring columns) issued up front
constexpr int NMS_RY = 4;

__global__ void __launch_bounds__(256)
nms_kernel(const float *__restrict__ eig, int H, int W, const uint8_t *__restrict__ mask, int64_t mask_pitch,
           double quality, GfttCounters *__restrict__ cnt, unsigned long long *__restrict__ keys, uint32_t cap)
{
    const uint32_t mb = cnt->maxbits;
    if (!mb) return;
    __shared__ uint32_t shist[4096];                              // per-CTA response histogram (top-k prefilter)
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) shist[i] = 0;
    __syncthreads();
    const float thr = (float)((double)dec_f32(mb) * quality);
    const int wq = (W + 3) >> 2;                                  // 4-pixel groups per row
    const int hq = (H - 2 + NMS_RY - 1) / NMS_RY;                 // row groups over rows 1 .. H-2
    const int64_t total = (int64_t)hq * wq;
    const int lane = threadIdx.x & 31;
    const bool vec = (W & 3) == 0;                                // rows are 16-byte aligned: one float4 + two scalars per row
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < total; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = base + threadIdx.x;
        uint32_t flags = 0;                                       // bit (4 * row + col)
        float r[NMS_RY + 2][6];                                   // rows y0-1 .. y0+RY, columns x0-1 .. x0+4, thresholded
        int y0 = 0, x0 = 0;
        if (g < total) {
            const int gy = (int)(g / wq);
            y0 = gy * NMS_RY + 1;
            x0 = (int)(g - (int64_t)gy * wq) * 4;
#pragma unroll
            for (int dy = 0; dy < NMS_RY + 2; dy++) {
                const int yy = min(y0 - 1 + dy, H - 1);
                const float *row = eig + (int64_t)yy * W;
                if (vec) {
                    const float4 m = __ldg(reinterpret_cast<const float4 *>(row + x0));
                    r[dy][0] = x0 > 0 ? __ldg(row + x0 - 1) : 0.f;
                    r[dy][1] = m.x; r[dy][2] = m.y; r[dy][3] = m.z; r[dy][4] = m.w;
                    r[dy][5] = x0 + 4 < W ? __ldg(row + x0 + 4) : 0.f;
                } else {
#pragma unroll
                    for (int c = 0; c < 6; c++) {
                        const int xx = x0 - 1 + c;
                        r[dy][c] = (xx >= 0 && xx < W) ? __ldg(row + xx) : 0.f;
                    }
                }
            }
#pragma unroll
            for (int dy = 0; dy < NMS_RY + 2; dy++)
#pragma unroll
                for (int c = 0; c < 6; c++) r[dy][c] = tozero(r[dy][c], thr);
#pragma unroll
            for (int ry = 0; ry < NMS_RY; ry++) {
                const int y = y0 + ry;
                if (y > H - 2) break;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int xx = x0 + i;
                    const float v = r[ry + 1][i + 1];
                    if (xx >= 1 && xx <= W - 2 && v != 0.f) {
                        float m = v;
#pragma unroll
                        for (int dy = 0; dy < 3; dy++)
                            m = fmaxf(m, fmaxf(r[ry + dy][i], fmaxf(r[ry + dy][i + 1], r[ry + dy][i + 2])));
                        if (v == m && (!mask || mask[(int64_t)y * mask_pitch + xx])) flags |= 1u << (4 * ry + i);
                    }
                }
            }
        }
        const int n = __popc(flags);
        // warp-aggregated append
        int pre = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int tmp = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += tmp; }
        const int wtot = __shfl_sync(0xffffffffu, pre, 31);
        if (wtot) {
            uint32_t basepos = 0;
            if (lane == 31) basepos = atomicAdd(&cnt->ncand, (uint32_t)wtot);
            basepos = __shfl_sync(0xffffffffu, basepos, 31);
            uint32_t pos = basepos + pre - n;
#pragma unroll
            for (int ry = 0; ry < NMS_RY; ry++)
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (flags & (1u << (4 * ry + i))) {
                        const uint32_t e = enc_f32(r[ry + 1][i + 1]);
                        if (pos < cap) keys[pos] = ((unsigned long long)e << 32) | (uint32_t)((y0 + ry) * W + x0 + i);
                        atomicAdd(&shist[e >> 20], 1u);
                        pos++;
                    }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4096; i += blockDim.x)
        if (shist[i]) atomicAdd(&cnt->hist[i], shist[i]);
}

// ---------------------------------------------------------------------------------------------
// Radix sort (ascending, 64-bit keys, stable LSD, 8-bit digits).
constexpr int RS_WARPS = 8, RS_STEPS = 16, RS_TILE = RS_WARPS * RS_STEPS * 32;   // 4096 keys per block

__global__ void __launch_bounds__(256)
complement_kernel(unsigned long long *__restrict__ keys, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ~keys[i];                   // ascending sort of ~key = descending (response, address)
}

// keep the candidates whose response bin is >= min_bin (the strongest ones), complemented for the ascending sort
__global__ void __launch_bounds__(256)
select_kernel(const unsigned long long *__restrict__ keys, uint32_t n, uint32_t min_bin, unsigned long long *__restrict__ out,
              uint32_t *__restrict__ out_count)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long k = 0;
    bool keep = false;
    if (i < n) { k = keys[i]; keep = (uint32_t)(k >> 52) >= min_bin; }
    const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
    if (ballot) {
        const int lane = threadIdx.x & 31;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(out_count, (uint32_t)__popc(ballot));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) out[base + __popc(ballot & ((1u << lane) - 1))] = ~k;
    }
}

__global__ void __launch_bounds__(256)
rs_count_kernel(const unsigned long long *__restrict__ keys, uint32_t n, int shift, uint32_t nblocks,
                uint32_t *__restrict__ blockhist /*[256][nblocks]*/)
{
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
    for (uint32_t i = threadIdx.x; i < RS_TILE; i += 256) {
        const uint32_t g = base + i;
        if (g < n) atomicAdd(&sh[(uint32_t)((keys[g] >> shift) & 0xff)], 1u);
    }
    __syncthreads();
    blockhist[threadIdx.x * nblocks + blockIdx.x] = sh[threadIdx.x];
}

// exclusive scan of uint32 in place.  Block b scans its SCAN_TILE elements; with `block_sums` the block total goes there
// (phase 1 of the hierarchical scan), with `block_offs` the scanned totals are added back (phase 3).
constexpr int SCAN_EPT = 8, SCAN_TILE = 1024 * SCAN_EPT;

__global__ void __launch_bounds__(1024)
scan_tile_kernel(uint32_t *__restrict__ data, uint32_t len, uint32_t *__restrict__ block_sums, uint32_t *__restrict__ total_out)
{
    __shared__ uint32_t warp_tot[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t i0 = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_EPT;
    uint32_t v[SCAN_EPT], sum = 0;
#pragma unroll
    for (int e = 0; e < SCAN_EPT; e++) { v[e] = (i0 + e < len) ? data[i0 + e] : 0; sum += v[e]; }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t tmp = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += tmp;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t tmp = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += tmp;
        }
        warp_tot[lane] = w;              // inclusive over warps
    }
    __syncthreads();
    uint32_t run = (wid ? warp_tot[wid - 1] : 0) + inc - sum;
#pragma unroll
    for (int e = 0; e < SCAN_EPT; e++) { if (i0 + e < len) data[i0 + e] = run; run += v[e]; }
    if (threadIdx.x == 1023) {
        if (block_sums) block_sums[blockIdx.x] = run;
        if (total_out) *total_out = run;
    }
}

__global__ void __launch_bounds__(1024)
scan_add_kernel(uint32_t *__restrict__ data, uint32_t len, const uint32_t *__restrict__ block_offs, uint32_t nblocks,
                uint32_t *__restrict__ total_out)
{
    const uint32_t off = block_offs[blockIdx.x];
    const uint32_t i0 = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_EPT;
#pragma unroll
    for (int e = 0; e < SCAN_EPT; e++) if (i0 + e < len) data[i0 + e] += off;
    // grand total = offset of the last block + its own sum, which phase 1 left in block_offs[nblocks] (see scan_u32)
    if (total_out && blockIdx.x == 0 && threadIdx.x == 0) *total_out = block_offs[nblocks];
}

// exclusive scan of data[0..len) in place; *total_out (device, optional) receives the grand total.
// scratch: (len / SCAN_TILE + 2) uint32.  len up to SCAN_TILE^2 (67 M).
static void scan_u32(uint32_t *data, uint32_t len, uint32_t *total_out, uint32_t *scratch, cudaStream_t st)
{
    const uint32_t nb = (len + SCAN_TILE - 1) / SCAN_TILE;
    if (nb <= 1) {
        scan_tile_kernel<<<1, 1024, 0, st>>>(data, len, nullptr, total_out);
        return;
    }
    scan_tile_kernel<<<nb, 1024, 0, st>>>(data, len, scratch, nullptr);
    // scan the nb block sums; the total lands in scratch[nb]
    scan_tile_kernel<<<1, 1024, 0, st>>>(scratch, nb, nullptr, scratch + nb);
    scan_add_kernel<<<nb, 1024, 0, st>>>(data, len, scratch, nb, total_out);
}

__global__ void __launch_bounds__(256)
rs_scatter_kernel(const unsigned long long *__restrict__ src, unsigned long long *__restrict__ dst, uint32_t n,
                  int shift, uint32_t nblocks, const uint32_t *__restrict__ blockoff /*[256][nblocks] scanned*/)
{
    __shared__ uint32_t wcount[RS_WARPS][256];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += 256) (&wcount[0][0])[i] = 0;
    __syncthreads();
    const uint32_t wbase = blockIdx.x * RS_TILE + wid * (RS_STEPS * 32);
    unsigned long long k[RS_STEPS];
    uint32_t rank[RS_STEPS];
#pragma unroll
    for (int s = 0; s < RS_STEPS; s++) {
        const uint32_t g = wbase + s * 32 + lane;
        const bool valid = g < n;
        k[s] = valid ? src[g] : 0ull;
        const uint32_t d = (uint32_t)((k[s] >> shift) & 0xff);
        const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
        uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : 0x100u + lane) & vmask;
        if (!valid) peers = 0;
        const uint32_t before = __popc(peers & ((1u << lane) - 1));
        uint32_t prev = 0;
        if (valid) prev = wcount[wid][d];          // all peers read the same value before the leader updates
        __syncwarp();
        if (valid && before == 0) wcount[wid][d] = prev + __popc(peers);
        __syncwarp();
        rank[s] = prev + before;                  // rank within this warp's chunk for digit d
    }
    __syncthreads();
    // exclusive prefix over warps per digit, plus the global offset of (digit, block)
    {
        const int d = threadIdx.x;
        uint32_t run = blockoff[d * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) { const uint32_t c = wcount[w][d]; wcount[w][d] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < RS_STEPS; s++) {
        const uint32_t g = wbase + s * 32 + lane;
        if (g < n) {
            const uint32_t d = (uint32_t)((k[s] >> shift) & 0xff);
            dst[wcount[wid][d] + rank[s]] = k[s];
        }
    }
}

// Stable LSD radix sort of n 64-bit keys on bytes [first_byte, last_byte] (shared with grid.cu).  keys0 holds the input,
// keys1 is the ping-pong buffer; returns the buffer that holds the result.  scratch: radix_sort_scratch_words(n) uint32.
size_t radix_sort_scratch_words(uint32_t n)
{
    const size_t nblocks = ((size_t)n + RS_TILE - 1) / RS_TILE;
    return 256 * nblocks + (256 * nblocks) / SCAN_TILE + 4;
}
unsigned long long *radix_sort_u64_bytes(unsigned long long *keys0, unsigned long long *keys1, uint32_t n, int first_byte,
                                         int last_byte, uint32_t *scratch, cudaStream_t st)
{
    if (n <= 1) return keys0;
    const uint32_t nblocks = (n + RS_TILE - 1) / RS_TILE;
    uint32_t *blockhist = scratch, *scan_scratch = scratch + 256 * (size_t)nblocks;
    unsigned long long *src = keys0, *dst = keys1;
    for (int p = first_byte; p <= last_byte; p++) {
        rs_count_kernel<<<nblocks, 256, 0, st>>>(src, n, 8 * p, nblocks, blockhist);
        scan_u32(blockhist, 256u * nblocks, nullptr, scan_scratch, st);
        rs_scatter_kernel<<<nblocks, 256, 0, st>>>(src, dst, n, 8 * p, nblocks, blockhist);
        unsigned long long *t = src; src = dst; dst = t;
    }
    return src;
}

// After the stable sort on the response bytes, equal responses keep their arrival order: order each run of equal
// responses by address (complemented keys ascending = address descending, OpenCV's tie-break).  Runs are rare and short;
// the first element of a run sorts it.
__global__ void __launch_bounds__(256)
tie_fix_kernel(unsigned long long *__restrict__ keys, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t v = (uint32_t)(keys[i] >> 32);
    if (i > 0 && (uint32_t)(keys[i - 1] >> 32) == v) return;       // not the first of its run
    uint32_t e = i + 1;
    while (e < n && (uint32_t)(keys[e] >> 32) == v) e++;
    for (uint32_t a = i + 1; a < e; a++) {                          // insertion sort of keys[i..e)
        const unsigned long long k = keys[a];
        uint32_t b = a;
        while (b > i && keys[b - 1] > k) { keys[b] = keys[b - 1]; b--; }
        keys[b] = k;
    }
}

// ---------------------------------------------------------------------------------------------
// K2c: cell grid + culling rounds.  sorted[r] = ~key of rank r (rank 0 = strongest).
__global__ void __launch_bounds__(256)
cell_count_kernel(const unsigned long long *__restrict__ sorted, uint32_t n, int W, int cell, int gw,
                  uint32_t *__restrict__ pos, uint32_t *__restrict__ cell_count, uint8_t *__restrict__ state)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t idx = (uint32_t)(~sorted[r]);
    const uint32_t y = idx / W, x = idx - y * W;
    pos[r] = x | (y << 16);
    state[r] = ST_UNDECIDED;
    atomicAdd(&cell_count[(y / cell) * gw + x / cell], 1u);
}

__global__ void __launch_bounds__(256)
cell_fill_kernel(const uint32_t *__restrict__ pos, uint32_t n, int cell, int gw, const uint32_t *__restrict__ cell_start,
                 uint32_t *__restrict__ cell_fill, uint32_t *__restrict__ items)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t p = pos[r];
    const uint32_t c = ((p >> 16) / cell) * gw + (p & 0xffffu) / cell;
    items[cell_start[c] + atomicAdd(&cell_fill[c], 1u)] = r;
}

__global__ void __launch_bounds__(256)
cull_round_kernel(const uint32_t *__restrict__ pos, uint32_t n, const uint32_t *__restrict__ cell_start,
                  const uint32_t *__restrict__ items, volatile uint8_t *state, GfttCounters *__restrict__ cnt,
                  int cell, int gw, int gh, float md2, int round)
{
    if (round > 0 && cnt->remaining[(round - 1) & 63] == 0) return;      // already converged
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    bool left = false;
    if (r < n && state[r] == ST_UNDECIDED) {
        const uint32_t p = pos[r];
        const int x = (int)(p & 0xffffu), y = (int)(p >> 16);
        const int cx = x / cell, cy = y / cell;
        const int x1 = max(cx - 1, 0), x2 = min(cx + 1, gw - 1), y1 = max(cy - 1, 0), y2 = min(cy + 1, gh - 1);
        bool blocked = false, killed = false;
        for (int yy = y1; yy <= y2 && !killed; yy++) {
            // the cells x1..x2 of one grid row are contiguous in the CSR layout
            const uint32_t kb = cell_start[yy * gw + x1], ke = cell_start[yy * gw + x2 + 1];
            for (uint32_t k = kb; k < ke; k++) {
                const uint32_t q = items[k];
                if (q >= r) continue;                                    // only stronger candidates matter
                const uint32_t pq = pos[q];
                const float dx = (float)(x - (int)(pq & 0xffffu)), dy = (float)(y - (int)(pq >> 16));
                if (dx * dx + dy * dy < md2) {
                    const uint8_t s = state[q];
                    if (s == ST_ACCEPTED) { killed = true; break; }
                    if (s == ST_UNDECIDED) blocked = true;
                }
            }
        }
        if (killed) state[r] = ST_REJECTED;
        else if (!blocked) state[r] = ST_ACCEPTED;
        else left = true;
    }
    const uint32_t ballot = __ballot_sync(0xffffffffu, left);
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(&cnt->remaining[round & 63], (uint32_t)__popc(ballot));
}

// accepted candidates in rank order: per-block counts -> scan -> scatter of (x, y)
__global__ void __launch_bounds__(256)
accepted_count_kernel(const uint8_t *__restrict__ state, uint32_t n, uint32_t *__restrict__ block_counts)
{
    const uint32_t r = blockIdx.x * 256 + threadIdx.x;
    const int c = __syncthreads_count(r < n && state[r] == ST_ACCEPTED);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = (uint32_t)c;
}

__global__ void __launch_bounds__(256)
write_corners_kernel(const uint32_t *__restrict__ pos, const uint8_t *__restrict__ state, uint32_t n,
                     const uint32_t *__restrict__ block_off, uint32_t limit, float *__restrict__ out_xy)
{
    __shared__ uint32_t wbase[8];
    const uint32_t r = blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool a = r < n && (!state || state[r] == ST_ACCEPTED);
    const uint32_t ballot = __ballot_sync(0xffffffffu, a);
    if (lane == 0) wbase[wid] = __popc(ballot);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = block_off ? block_off[blockIdx.x] : blockIdx.x * 256;
        for (int w = 0; w < 8; w++) { const uint32_t c = wbase[w]; wbase[w] = run; run += c; }
    }
    __syncthreads();
    if (!a) return;
    const uint32_t m = wbase[wid] + __popc(ballot & ((1u << lane) - 1));
    if (m >= limit) return;
    const uint32_t p = pos[r];
    out_xy[2 * m] = (float)(p & 0xffffu);
    out_xy[2 * m + 1] = (float)(p >> 16);
}

// workspace layout
struct GfttLayout {
    size_t off_cnt, off_eig, off_keys0, off_keys1, off_blockhist, off_pos, off_state, off_cells, off_fill, off_items,
        off_blockcnt, off_scan, total;
    uint32_t cap, max_sort_blocks, max_cells;
};
static GfttLayout gftt_layout(int H, int W)
{
    GfttLayout L;
    const size_t np = (size_t)H * W;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    L.cap = (uint32_t)(np / 4 + 4096);                 // 3x3 NMS maxima cannot be denser than 1 in 4 pixels
    L.max_sort_blocks = (L.cap + RS_TILE - 1) / RS_TILE;
    L.max_cells = (uint32_t)np + 1;                    // cell >= 1 pixel
    size_t o = 0;
    L.off_cnt = o; o = up(o + sizeof(GfttCounters));
    L.off_eig = o; o = up(o + np * 4);
    L.off_keys0 = o; o = up(o + (size_t)L.cap * 8);
    L.off_keys1 = o; o = up(o + (size_t)L.cap * 8);
    L.off_blockhist = o; o = up(o + (size_t)256 * L.max_sort_blocks * 4);
    L.off_pos = o; o = up(o + (size_t)L.cap * 4);
    L.off_state = o; o = up(o + (size_t)L.cap);
    L.off_cells = o; o = up(o + ((size_t)L.max_cells + 1) * 4);
    L.off_fill = o; o = up(o + (size_t)L.max_cells * 4);
    L.off_items = o; o = up(o + (size_t)L.cap * 4);
    L.off_blockcnt = o; o = up(o + ((size_t)L.cap / 256 + 2) * 4);
    L.off_scan = o; o = up(o + ((size_t)L.max_cells / SCAN_TILE + 4) * 4);
    L.total = o;
    return L;
}

} // namespace ibt

IBT_API int ibt_min_eigen_f32(const uint8_t *gray, int H, int W, int64_t pitch, int blockSize, float *eig,
                              int64_t eig_pitch, void *stream)
{
    return ibt::launch_eig(gray, H, W, pitch, blockSize, eig, eig_pitch, nullptr, 0, nullptr, static_cast<cudaStream_t>(stream));
}

IBT_API size_t ibt_gftt_workspace_bytes(int H, int W)
{
    if (H <= 0 || W <= 0) return 0;
    return ibt::gftt_layout(H, W).total;
}

IBT_API int ibt_gftt(const uint8_t *gray, int64_t pitch, const uint8_t *mask, int64_t mask_pitch, int H, int W,
                     int maxCorners, double qualityLevel, double minDistance, int blockSize, void *workspace,
                     size_t workspace_bytes, float *out_xy, int cap, int *out_count, void *stream)
{
    using namespace ibt;
    if (!gray || !workspace || !out_count || H < 3 || W < 3 || H > 65535 || W > 65535 || (int64_t)H * W > 0x7fffffffLL ||
        cap < 0 || (cap > 0 && !out_xy) || (mask && mask_pitch < W) || qualityLevel < 0 || minDistance < 0 || minDistance > 1024)
        return IBT_E_INVALID;
    const GfttLayout L = gftt_layout(H, W);
    if (workspace_bytes < L.total) return IBT_E_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    GfttCounters *cnt = reinterpret_cast<GfttCounters *>(ws + L.off_cnt);
    float *eig = reinterpret_cast<float *>(ws + L.off_eig);
    unsigned long long *keys0 = reinterpret_cast<unsigned long long *>(ws + L.off_keys0);
    unsigned long long *keys1 = reinterpret_cast<unsigned long long *>(ws + L.off_keys1);
    uint32_t *blockhist = reinterpret_cast<uint32_t *>(ws + L.off_blockhist);
    uint32_t *pos = reinterpret_cast<uint32_t *>(ws + L.off_pos);
    uint8_t *state = ws + L.off_state;
    uint32_t *cell_start = reinterpret_cast<uint32_t *>(ws + L.off_cells);
    uint32_t *cell_fill = reinterpret_cast<uint32_t *>(ws + L.off_fill);
    uint32_t *items = reinterpret_cast<uint32_t *>(ws + L.off_items);
    uint32_t *blockcnt = reinterpret_cast<uint32_t *>(ws + L.off_blockcnt);
    uint32_t *scan_scratch = reinterpret_cast<uint32_t *>(ws + L.off_scan);
    *out_count = 0;

    IBT_CUDA_TRY(cudaMemsetAsync(cnt, 0, sizeof(GfttCounters), st));
    int rc = launch_eig(gray, H, W, pitch, blockSize, eig, (int64_t)W * 4, mask, mask_pitch, &cnt->maxbits, st);
    if (rc) return rc;
    const int nblk = kNumSMs * 8;
    const bool cull = minDistance >= 1.0;
    const int cell = cull ? (int)lrint(minDistance) : 65536;        // no culling: one cell, only positions are needed
    const int gw = (W + cell - 1) / cell, gh = (H + cell - 1) / cell;
    const uint32_t ncells = (uint32_t)gw * gh;
    uint32_t limit = (maxCorners > 0) ? (uint32_t)maxCorners : 0xffffffffu;
    if (limit > (uint32_t)cap) limit = (uint32_t)cap;
    static thread_local GfttCounters hc;                           // 17 KB: keep it off the stack
    uint32_t nout = 0;

    // attempt 0 may work on the strongest candidates only (top-k prefilter); if culling leaves fewer than maxCorners of
    // them, attempt 1 repeats with every candidate.  The greedy order makes the prefix exact either way: whether a
    // candidate is accepted depends on stronger candidates only.
    for (int attempt = 0; attempt < 2; attempt++) {
        if (attempt == 1) {
            IBT_CUDA_TRY(cudaMemsetAsync(&cnt->ncand, 0, sizeof(GfttCounters) - offsetof(GfttCounters, ncand), st));
        }
        nms_kernel<<<nblk, 256, 0, st>>>(eig, H, W, mask, mask_pitch, qualityLevel, cnt, keys0, L.cap);
        if ((rc = check_launch("nms_kernel"))) return rc;
        IBT_CUDA_TRY(cudaMemcpyAsync(&hc, cnt, sizeof(hc), cudaMemcpyDeviceToHost, st));
        IBT_CUDA_TRY(cudaStreamSynchronize(st));
        if (hc.ncand > L.cap) return IBT_E_CAPACITY;
        uint32_t n = hc.ncand;
        if (n == 0) return IBT_OK;

        // ---- order the candidates by (response desc, address desc): rank ------------------------------------
        unsigned long long *src = keys0, *dst = keys1;
        bool subset = false;
        const uint64_t want = maxCorners > 0 ? (uint64_t)maxCorners * 4 + 1024 : 0;
        if (attempt == 0 && maxCorners > 0 && cull && n > 2 * want) {
            uint32_t cum = 0;
            int b = 4095;
            for (; b >= 0; b--) { cum += hc.hist[b]; if (cum >= want) break; }
            if (b > 0 && cum < n) {
                IBT_CUDA_TRY(cudaMemsetAsync(&cnt->nacc, 0, 4, st));            // reused as the selection counter
                select_kernel<<<(n + 255) / 256, 256, 0, st>>>(keys0, n, (uint32_t)b, keys1, &cnt->nacc);
                src = keys1; dst = keys0;
                n = cum;
                subset = true;
            }
        }
        if (!subset) complement_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, n);
        if (n > 1) {
            const uint32_t nblocks = (n + RS_TILE - 1) / RS_TILE;
            // the top response byte (sign + 7 exponent bits) is usually the same for every candidate: the histogram tells
            int lo_bin = 4096, hi_bin = -1;
            for (int b = 0; b < 4096; b++) if (hc.hist[b]) { if (b < lo_bin) lo_bin = b; hi_bin = b; }
            const int last_pass = (lo_bin >> 4) == (hi_bin >> 4) ? 6 : 7;
            for (int p = 4; p <= last_pass; p++) {                      // response bytes only; ties are fixed below
                rs_count_kernel<<<nblocks, 256, 0, st>>>(src, n, 8 * p, nblocks, blockhist);
                scan_u32(blockhist, 256u * nblocks, nullptr, scan_scratch, st);
                rs_scatter_kernel<<<nblocks, 256, 0, st>>>(src, dst, n, 8 * p, nblocks, blockhist);
                unsigned long long *t = src; src = dst; dst = t;
            }
            tie_fix_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, n);
            if ((rc = check_launch("radix sort"))) return rc;
        }
        const uint32_t nb = (n + 255) / 256;

        // positions of all ranks (+ cell histogram when culling)
        IBT_CUDA_TRY(cudaMemsetAsync(cell_start, 0, ((size_t)ncells + 1) * 4, st));
        cell_count_kernel<<<nb, 256, 0, st>>>(src, n, W, cell, gw, pos, cell_start, state);
        if ((rc = check_launch("cell_count_kernel"))) return rc;
        nout = n;
        if (cull) {
            IBT_CUDA_TRY(cudaMemsetAsync(cell_fill, 0, (size_t)ncells * 4, st));
            scan_u32(cell_start, ncells + 1, nullptr, scan_scratch, st);
            cell_fill_kernel<<<nb, 256, 0, st>>>(pos, n, cell, gw, cell_start, cell_fill, items);
            const float md2 = (float)(minDistance * minDistance);
            int round = 0;
            for (;;) {
                const int batch = round == 0 ? 12 : 4;
                for (int b = 0; b < batch; b++, round++) {
                    if (round >= 64) IBT_CUDA_TRY(cudaMemsetAsync(&cnt->remaining[round & 63], 0, 4, st));
                    cull_round_kernel<<<nb, 256, 0, st>>>(pos, n, cell_start, items, state, cnt, cell, gw, gh, md2, round);
                }
                if ((rc = check_launch("cull_round_kernel"))) return rc;
                IBT_CUDA_TRY(cudaMemcpyAsync(&hc, cnt, offsetof(GfttCounters, hist), cudaMemcpyDeviceToHost, st));
                IBT_CUDA_TRY(cudaStreamSynchronize(st));
                if (hc.remaining[(round - 1) & 63] == 0) break;
                if (round > (1 << 20)) return IBT_E_CUDA;
            }
            accepted_count_kernel<<<nb, 256, 0, st>>>(state, n, blockcnt);
            scan_u32(blockcnt, nb, &cnt->nacc, scan_scratch, st);
            write_corners_kernel<<<nb, 256, 0, st>>>(pos, state, n, blockcnt, limit, out_xy);
            if ((rc = check_launch("write_corners_kernel"))) return rc;
            IBT_CUDA_TRY(cudaMemcpyAsync(&hc, cnt, offsetof(GfttCounters, hist), cudaMemcpyDeviceToHost, st));
            IBT_CUDA_TRY(cudaStreamSynchronize(st));
            nout = hc.nacc;
        } else {
            write_corners_kernel<<<nb, 256, 0, st>>>(pos, nullptr, n, nullptr, limit, out_xy);
            if ((rc = check_launch("write_corners_kernel"))) return rc;
        }
        if (!subset || nout >= (uint32_t)maxCorners) break;          // done (the subset produced a full prefix)
    }
    if (maxCorners > 0 && nout > (uint32_t)maxCorners) nout = (uint32_t)maxCorners;
    *out_count = (int)nout;
    return nout > (uint32_t)cap ? IBT_E_CAPACITY : IBT_OK;
}
