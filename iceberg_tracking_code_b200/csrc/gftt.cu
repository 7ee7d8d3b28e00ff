// K2: Shi-Tomasi corner seeding.  Replaces
// cv2.goodFeaturesToTrack(frame_gray, mask=mask, **feature_params), s1_lucaskanade_tracking.py:437
// (s0_1_test_lucaskanade_tracking.py:167).
//
//   K2a  eig_kernel      Sobel3 -> (dx^2, dxdy, dy^2) float -> blockSize^2 box sum (double) ->
//                        lambda_min map + masked global max (ordered-uint atomicMax)
//   K2b  nms_kernel      threshold (q * max), 3x3 non-maximum suppression, mask, 1-px border;
//                        candidates appended as 64-bit keys (response bits << 32 | linear address)
//                        and marked UNDECIDED in a per-pixel state map
//   K2c  cull_round      OpenCV's sequential greedy min-distance culling restated as a fixed point
//                        that is safe to run in parallel: a candidate is REJECTED once a stronger
//                        conflicting candidate is ACCEPTED, ACCEPTED once all stronger conflicting
//                        candidates are REJECTED ("conflicting" = closer than minDistance AND in the
//                        3x3 cell neighbourhood of OpenCV's bucket grid).  Decisions are final and
//                        monotone, so rounds may read each other's fresh or stale states.
//        radix sort      stable LSD, 8-bit digits, warp match_any ranking; orders accepted keys by
//                        (response desc, address desc) exactly like OpenCV's comparator
#include "common.cuh"
#include <string.h>

namespace ibt {

// ---------------------------------------------------------------------------------------------
// K2a
constexpr int ETW = 64, ETH = 16;
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

// products at product-image position (py, px), both already inside the image
__device__ __forceinline__ void sobel_products(const uint8_t *__restrict__ img, int64_t pitch, int H, int W,
                                               int py, int px, float scale, float &xx, float &xy, float &yy)
{
    const uint8_t *r0 = img + (int64_t)r101(py - 1, H) * pitch;
    const uint8_t *r1 = img + (int64_t)py * pitch;
    const uint8_t *r2 = img + (int64_t)r101(py + 1, H) * pitch;
    const int xm = r101(px - 1, W), xp = r101(px + 1, W);
    const int sx = (r0[xp] + 2 * r1[xp] + r2[xp]) - (r0[xm] + 2 * r1[xm] + r2[xm]);
    const int sy = (r2[xm] + 2 * r2[px] + r2[xp]) - (r0[xm] + 2 * r0[px] + r0[xp]);
    const float dx = __fmul_rn((float)sx, scale), dy = __fmul_rn((float)sy, scale);
    xx = __fmul_rn(dx, dx); xy = __fmul_rn(dx, dy); yy = __fmul_rn(dy, dy);
}

__global__ void __launch_bounds__(256)
eig_kernel(const uint8_t *__restrict__ img, int H, int W, int64_t pitch, int bs, float scale,
           float *__restrict__ eig, int64_t eig_pitch_f,
           const uint8_t *__restrict__ mask, int64_t mask_pitch, uint32_t *__restrict__ maxbits)
{
    extern __shared__ __align__(16) unsigned char eig_smem[];
    const int PR = ETH + bs - 1, PC = ETW + bs - 1, PCS = PC + 1;
    float *prod = reinterpret_cast<float *>(eig_smem);                         // [3][PR][PCS]
    double *hsum = reinterpret_cast<double *>(eig_smem + (((size_t)3 * PR * PCS * 4 + 15) & ~(size_t)15));   // [3][PR][ETW]
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * ETW, y0 = blockIdx.y * ETH;
    const int a0 = -(bs / 2);

    // phase 0: products over the (PR x PC) support; REFLECT_101 acts on the product image
    for (int idx = tid; idx < PR * PC; idx += 256) {
        const int r = idx / PC, c = idx - r * PC;
        const int py = r101(y0 + a0 + r, H), px = r101(x0 + a0 + c, W);
        float xx, xy, yy;
        sobel_products(img, pitch, H, W, py, px, scale, xx, xy, yy);
        prod[(0 * PR + r) * PCS + c] = xx;
        prod[(1 * PR + r) * PCS + c] = xy;
        prod[(2 * PR + r) * PCS + c] = yy;
    }
    __syncthreads();
    // phase A: horizontal sums in double
    for (int idx = tid; idx < 3 * PR * ETW; idx += 256) {
        const int x = idx % ETW, pr = idx / ETW;                 // pr = plane * PR + r
        const float *p = prod + pr * PCS + x;
        double s = 0.0;
        for (int j = 0; j < bs; j++) s += (double)p[j];
        hsum[pr * ETW + x] = s;
    }
    __syncthreads();
    // phase B: vertical sums, lambda_min, masked max
    uint32_t lmax = 0;
    for (int idx = tid; idx < ETH * ETW; idx += 256) {
        const int x = idx % ETW, y = idx / ETW;
        const int gx = x0 + x, gy = y0 + y;
        if (gx >= W || gy >= H) continue;
        double sxx = 0.0, sxy = 0.0, syy = 0.0;
        for (int i = 0; i < bs; i++) {
            sxx += hsum[(0 * PR + y + i) * ETW + x];
            sxy += hsum[(1 * PR + y + i) * ETW + x];
            syy += hsum[(2 * PR + y + i) * ETW + x];
        }
        const float a = __fmul_rn((float)sxx, 0.5f), b = (float)sxy, c = __fmul_rn((float)syy, 0.5f);
        const float d = __fsub_rn(a, c);
        const float e = __fsub_rn(__fadd_rn(a, c), __fsqrt_rn(__fadd_rn(__fmul_rn(d, d), __fmul_rn(b, b))));
        eig[(int64_t)gy * eig_pitch_f + gx] = e;
        if (maxbits && (!mask || mask[(int64_t)gy * mask_pitch + gx])) lmax = max(lmax, enc_f32(e));
    }
    if (maxbits) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        if ((tid & 31) == 0 && lmax) atomicMax(maxbits, lmax);
    }
}

static size_t eig_smem_bytes(int bs)
{
    const int PR = ETH + bs - 1, PC = ETW + bs - 1, PCS = PC + 1;
    size_t a = ((size_t)3 * PR * PCS * 4 + 15) & ~(size_t)15;
    return a + (size_t)3 * PR * ETW * 8;
}

static int launch_eig(const uint8_t *gray, int H, int W, int64_t pitch, int bs, float *eig, int64_t eig_pitch_bytes,
                      const uint8_t *mask, int64_t mask_pitch, uint32_t *maxbits, cudaStream_t st)
{
    if (!gray || !eig || H <= 0 || W <= 0 || bs < 1 || bs > MAX_BLOCK || pitch < W || eig_pitch_bytes % 4 != 0 ||
        eig_pitch_bytes < (int64_t)W * 4)
        return IBT_E_INVALID;
    const size_t smem = eig_smem_bytes(bs);
    static bool attr_done = false;
    if (!attr_done) {
        IBT_CUDA_TRY(cudaFuncSetAttribute(eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eig_smem_bytes(MAX_BLOCK)));
        attr_done = true;
    }
    const float scale = (float)(1.0 / (4.0 * bs * 255.0));
    dim3 grid((W + ETW - 1) / ETW, (H + ETH - 1) / ETH);
    eig_kernel<<<grid, 256, smem, st>>>(gray, H, W, pitch, bs, scale, eig, eig_pitch_bytes / 4, mask, mask_pitch, maxbits);
    return check_launch("eig_kernel");
}

// ---------------------------------------------------------------------------------------------
// K2b
enum : uint8_t { ST_NONE = 0, ST_UNDECIDED = 1, ST_ACCEPTED = 2, ST_REJECTED = 3 };

struct GfttCounters {
    uint32_t maxbits;        // enc_f32 of the masked maximum, 0 = no allowed pixel
    uint32_t ncand;          // candidates found (may exceed capacity)
    uint32_t nacc;           // accepted after culling
    uint32_t pad;
    uint32_t remaining[64];  // undecided candidates left after round k (mod 64)
};

__device__ __forceinline__ float tozero(float v, float thr) { return v > thr ? v : 0.f; }

__global__ void __launch_bounds__(256)
nms_kernel(const float *__restrict__ eig, int H, int W, const uint8_t *__restrict__ mask, int64_t mask_pitch,
           double quality, GfttCounters *__restrict__ cnt, uint8_t *__restrict__ state,
           unsigned long long *__restrict__ keys, uint32_t cap, int cull)
{
    const uint32_t mb = cnt->maxbits;
    const float thr = mb ? (float)((double)dec_f32(mb) * quality) : 0.f;
    const int64_t total = (int64_t)H * W;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < total; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = base + threadIdx.x;
        bool is_cand = false;
        float v = 0.f;
        if (i < total && mb) {
            const int y = (int)(i / W), x = (int)(i - (int64_t)y * W);
            if (x >= 1 && x <= W - 2 && y >= 1 && y <= H - 2) {
                v = tozero(eig[i], thr);
                if (v != 0.f && (!mask || mask[(int64_t)y * mask_pitch + x])) {
                    float m = v;
#pragma unroll
                    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                        for (int dx = -1; dx <= 1; dx++) m = fmaxf(m, tozero(eig[i + dy * W + dx], thr));
                    is_cand = (v == m);
                }
            }
        }
        if (i < total) state[i] = is_cand ? (cull ? ST_UNDECIDED : ST_ACCEPTED) : ST_NONE;
        const uint32_t ballot = __ballot_sync(0xffffffffu, is_cand);
        if (ballot) {
            const int lane = threadIdx.x & 31;
            uint32_t basepos = 0;
            if (lane == 0) basepos = atomicAdd(&cnt->ncand, __popc(ballot));
            basepos = __shfl_sync(0xffffffffu, basepos, 0);
            if (is_cand) {
                const uint32_t pos = basepos + __popc(ballot & ((1u << lane) - 1));
                if (pos < cap) keys[pos] = ((unsigned long long)enc_f32(v) << 32) | (uint32_t)i;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2c: one round of the culling fixed point; one warp per candidate, lanes sweep dx.
__global__ void __launch_bounds__(256)
cull_round_kernel(const unsigned long long *__restrict__ keys, GfttCounters *__restrict__ cnt, uint32_t cap,
                  const float *__restrict__ eig, volatile uint8_t *state, int H, int W, double quality,
                  int R, float md2, int cell, int round)
{
    if (round > 0 && cnt->remaining[(round - 1) & 63] == 0) return;      // already converged
    const uint32_t n = min(cnt->ncand, cap);
    const uint32_t mb = cnt->maxbits;
    const float thr = mb ? (float)((double)dec_f32(mb) * quality) : 0.f;
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    uint32_t left = 0;
    for (uint32_t c = warp; c < n; c += nwarps) {
        const unsigned long long key = keys[c];
        const uint32_t idx = (uint32_t)key;
        if (state[idx] != ST_UNDECIDED) continue;
        const int y = idx / W, x = idx - y * W;
        const int cx = x / cell, cy = y / cell;
        bool blocked = false, killed = false;
        for (int dy = -R; dy <= R; dy++) {
            const int qy = y + dy;
            if (qy < 1 || qy > H - 2) continue;
            const int qcy = qy / cell;
            if (qcy < cy - 1 || qcy > cy + 1) continue;
            for (int dx0 = -R; dx0 <= R; dx0 += 32) {
                const int dx = dx0 + lane;
                const int qx = x + dx;
                if (dx <= R && qx >= 1 && qx <= W - 2 && (dx | dy) != 0) {
                    const float d2 = (float)(dx * dx + dy * dy);
                    const int qcx = qx / cell;
                    if (d2 < md2 && qcx >= cx - 1 && qcx <= cx + 1) {
                        const uint32_t q = (uint32_t)qy * W + qx;
                        const uint8_t s = state[q];
                        if (s == ST_UNDECIDED || s == ST_ACCEPTED) {
                            const unsigned long long qkey = ((unsigned long long)enc_f32(tozero(eig[q], thr)) << 32) | q;
                            if (qkey > key) { if (s == ST_ACCEPTED) killed = true; else blocked = true; }
                        }
                    }
                }
            }
        }
        killed = __any_sync(0xffffffffu, killed);
        blocked = __any_sync(0xffffffffu, blocked);
        if (lane == 0) {
            if (killed) state[idx] = ST_REJECTED;
            else if (!blocked) state[idx] = ST_ACCEPTED;
            else left++;
        }
    }
    if (lane == 0 && left) atomicAdd(&cnt->remaining[round & 63], left);
}

__global__ void __launch_bounds__(256)
collect_accepted_kernel(const unsigned long long *__restrict__ keys, GfttCounters *__restrict__ cnt, uint32_t cap,
                        const uint8_t *__restrict__ state, unsigned long long *__restrict__ out)
{
    const uint32_t n = min(cnt->ncand, cap);
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t c = base + threadIdx.x;
        unsigned long long key = 0;
        bool acc = false;
        if (c < n) { key = keys[c]; acc = state[(uint32_t)key] == ST_ACCEPTED; }
        const uint32_t ballot = __ballot_sync(0xffffffffu, acc);
        if (ballot) {
            const int lane = threadIdx.x & 31;
            uint32_t basepos = 0;
            if (lane == 0) basepos = atomicAdd(&cnt->nacc, __popc(ballot));
            basepos = __shfl_sync(0xffffffffu, basepos, 0);
            if (acc) out[basepos + __popc(ballot & ((1u << lane) - 1))] = ~key;   // complemented: ascending sort = descending key
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Radix sort (ascending, 64-bit keys, stable LSD, 8-bit digits).
constexpr int RS_WARPS = 8, RS_STEPS = 8, RS_TILE = RS_WARPS * RS_STEPS * 32;   // 2048 keys per block

__global__ void __launch_bounds__(256)
rs_hist_kernel(const unsigned long long *__restrict__ keys, uint32_t n, uint32_t *__restrict__ hist /*[8][256]*/)
{
    __shared__ uint32_t sh[8 * 256];
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long k = keys[i];
#pragma unroll
        for (int p = 0; p < 8; p++) atomicAdd(&sh[p * 256 + (uint32_t)((k >> (8 * p)) & 0xff)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
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

// exclusive scan of `len` uint32 in place, single block of 1024 threads
__global__ void __launch_bounds__(1024)
rs_scan_kernel(uint32_t *__restrict__ data, uint32_t len)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t base = 0; base < len; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < len ? data[i] : 0;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;              // inclusive over warps
        }
        __syncthreads();
        const uint32_t woff = wid ? warp_tot[wid - 1] : 0;
        const uint32_t c = carry;
        if (i < len) data[i] = c + woff + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + woff + inc;
        __syncthreads();
    }
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
        if (valid) {
            prev = wcount[wid][d];          // all peers read the same value before the leader updates
        }
        __syncwarp();
        if (valid && before == 0) wcount[wid][d] = prev + __popc(peers);
        __syncwarp();
        rank[s] = prev + before;            // rank within this warp's chunk for digit d
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

__global__ void __launch_bounds__(256)
write_corners_kernel(const unsigned long long *__restrict__ sorted, uint32_t nout, int W, float *__restrict__ out_xy)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nout) return;
    const uint32_t idx = (uint32_t)(~sorted[i]);
    const int y = idx / W, x = idx - y * W;
    out_xy[2 * i] = (float)x;
    out_xy[2 * i + 1] = (float)y;
}

// workspace layout
struct GfttLayout {
    size_t off_cnt, off_hist, off_eig, off_state, off_keys0, off_keys1, off_blockhist, total;
    uint32_t cap, max_blocks;
};
static GfttLayout gftt_layout(int H, int W)
{
    GfttLayout L;
    const size_t np = (size_t)H * W;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    L.cap = (uint32_t)(np / 4 + 4096);
    L.max_blocks = (L.cap + RS_TILE - 1) / RS_TILE;
    size_t o = 0;
    L.off_cnt = o; o = up(o + sizeof(GfttCounters));
    L.off_hist = o; o = up(o + 8 * 256 * 4);
    L.off_eig = o; o = up(o + np * 4);
    L.off_state = o; o = up(o + np);
    L.off_keys0 = o; o = up(o + (size_t)L.cap * 8);
    L.off_keys1 = o; o = up(o + (size_t)L.cap * 8);
    L.off_blockhist = o; o = up(o + (size_t)256 * L.max_blocks * 4);
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
    if (!gray || !workspace || !out_count || H < 3 || W < 3 || (int64_t)H * W > 0x7fffffffLL || cap < 0 || (cap > 0 && !out_xy) ||
        (mask && mask_pitch < W) || qualityLevel < 0 || minDistance < 0 || minDistance > 1024)
        return IBT_E_INVALID;
    const GfttLayout L = gftt_layout(H, W);
    if (workspace_bytes < L.total) return IBT_E_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    GfttCounters *cnt = reinterpret_cast<GfttCounters *>(ws + L.off_cnt);
    uint32_t *hist = reinterpret_cast<uint32_t *>(ws + L.off_hist);
    float *eig = reinterpret_cast<float *>(ws + L.off_eig);
    uint8_t *state = ws + L.off_state;
    unsigned long long *keys0 = reinterpret_cast<unsigned long long *>(ws + L.off_keys0);
    unsigned long long *keys1 = reinterpret_cast<unsigned long long *>(ws + L.off_keys1);
    uint32_t *blockhist = reinterpret_cast<uint32_t *>(ws + L.off_blockhist);
    *out_count = 0;

    IBT_CUDA_TRY(cudaMemsetAsync(cnt, 0, sizeof(GfttCounters), st));
    int rc = launch_eig(gray, H, W, pitch, blockSize, eig, (int64_t)W * 4, mask, mask_pitch, &cnt->maxbits, st);
    if (rc) return rc;
    const bool cull = minDistance >= 1.0;
    const int nblk = kNumSMs * 8;
    nms_kernel<<<nblk, 256, 0, st>>>(eig, H, W, mask, mask_pitch, qualityLevel, cnt, state, keys0, L.cap, cull ? 1 : 0);
    if ((rc = check_launch("nms_kernel"))) return rc;

    GfttCounters hc;
    if (cull) {
        const int cell = (int)lrint(minDistance);
        const int R = (int)ceil(minDistance) - 1;
        const float md2 = (float)(minDistance * minDistance);
        int round = 0;
        for (;;) {
            for (int b = 0; b < 16; b++, round++) {
                if (round >= 64) IBT_CUDA_TRY(cudaMemsetAsync(&cnt->remaining[round & 63], 0, 4, st));
                cull_round_kernel<<<nblk, 256, 0, st>>>(keys0, cnt, L.cap, eig, state, H, W, qualityLevel, R, md2, cell, round);
            }
            if ((rc = check_launch("cull_round_kernel"))) return rc;
            IBT_CUDA_TRY(cudaMemcpyAsync(&hc, cnt, sizeof(hc), cudaMemcpyDeviceToHost, st));
            IBT_CUDA_TRY(cudaStreamSynchronize(st));
            if (hc.ncand > L.cap) return IBT_E_CAPACITY;
            if (hc.ncand == 0 || hc.remaining[(round - 1) & 63] == 0) break;
            if (round > (1 << 20)) return IBT_E_CUDA;
        }
        if (hc.ncand == 0) return IBT_OK;
    } else {
        IBT_CUDA_TRY(cudaMemcpyAsync(&hc, cnt, sizeof(hc), cudaMemcpyDeviceToHost, st));
        IBT_CUDA_TRY(cudaStreamSynchronize(st));
        if (hc.ncand > L.cap) return IBT_E_CAPACITY;
        if (hc.ncand == 0) return IBT_OK;
    }
    collect_accepted_kernel<<<nblk, 256, 0, st>>>(keys0, cnt, L.cap, state, keys1);
    if ((rc = check_launch("collect_accepted_kernel"))) return rc;
    IBT_CUDA_TRY(cudaMemsetAsync(hist, 0, 8 * 256 * 4, st));
    // the accepted count is needed on the host to size the sort
    IBT_CUDA_TRY(cudaMemcpyAsync(&hc, cnt, sizeof(hc), cudaMemcpyDeviceToHost, st));
    IBT_CUDA_TRY(cudaStreamSynchronize(st));
    const uint32_t n = hc.nacc;
    if (n == 0) return IBT_OK;
    unsigned long long *src = keys1, *dst = keys0;
    if (n > 1) {
        uint32_t hhist[8 * 256];
        rs_hist_kernel<<<min((n + 255u) / 256u, (uint32_t)nblk), 256, 0, st>>>(src, n, hist);
        if ((rc = check_launch("rs_hist_kernel"))) return rc;
        IBT_CUDA_TRY(cudaMemcpyAsync(hhist, hist, sizeof(hhist), cudaMemcpyDeviceToHost, st));
        IBT_CUDA_TRY(cudaStreamSynchronize(st));
        const uint32_t nblocks = (n + RS_TILE - 1) / RS_TILE;
        for (int p = 0; p < 8; p++) {
            bool trivial = false;
            for (int d = 0; d < 256; d++) if (hhist[p * 256 + d] == n) { trivial = true; break; }
            if (trivial) continue;
            rs_count_kernel<<<nblocks, 256, 0, st>>>(src, n, 8 * p, nblocks, blockhist);
            rs_scan_kernel<<<1, 1024, 0, st>>>(blockhist, 256u * nblocks);
            rs_scatter_kernel<<<nblocks, 256, 0, st>>>(src, dst, n, 8 * p, nblocks, blockhist);
            unsigned long long *t = src; src = dst; dst = t;
        }
        if ((rc = check_launch("radix sort"))) return rc;
    }
    uint32_t nout = n;
    if (maxCorners > 0 && nout > (uint32_t)maxCorners) nout = (uint32_t)maxCorners;
    *out_count = (int)nout;
    const uint32_t nwrite = nout > (uint32_t)cap ? (uint32_t)cap : nout;
    if (nwrite) {
        write_corners_kernel<<<(nwrite + 255) / 256, 256, 0, st>>>(src, nwrite, W, out_xy);
        if ((rc = check_launch("write_corners_kernel"))) return rc;
    }
    IBT_CUDA_TRY(cudaStreamSynchronize(st));
    return nout > (uint32_t)cap ? IBT_E_CAPACITY : IBT_OK;
}
