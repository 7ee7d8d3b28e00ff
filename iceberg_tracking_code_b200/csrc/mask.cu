// Fjord-mask rasterisation: replaces Camera.mask_meshgrid (imports/camtools.py:184-211), which
// runs matplotlib.path.Path(poly).contains_points over all H*W pixel centres once per day folder
// (s1_lucaskanade_tracking.py:285-294).  Same crossing-number rule as matplotlib's point_in_path (the Graphics
// Gems "crossings" test, polygon implicitly closed, fp64), one thread per pixel, edges staged in shared memory.
// HBM-bound on the H*W byte store; the E edges are broadcast reads.
#include "common.cuh"

namespace ibt {

constexpr int MAX_POLY = 2048;

__global__ void __launch_bounds__(256)
polygon_mask_kernel(const double2 *__restrict__ poly, int E, int H, int W, uint8_t *__restrict__ out, int64_t pitch,
                    uint8_t inside_value)
{
    extern __shared__ double2 sp[];
    for (int i = threadIdx.x; i < E; i += blockDim.x) sp[i] = poly[i];
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W || y >= H) return;
    const double tx = (double)x, ty = (double)y;
    double2 v0 = sp[E - 1];
    bool yflag0 = v0.y >= ty;
    bool inside = false;
    for (int i = 0; i < E; i++) {
        const double2 v1 = sp[i];
        const bool yflag1 = v1.y >= ty;
        if (yflag0 != yflag1) {
            if (((v1.y - ty) * (v0.x - v1.x) >= (v1.x - tx) * (v0.y - v1.y)) == yflag1) inside = !inside;
        }
        yflag0 = yflag1;
        v0 = v1;
    }
    out[(int64_t)y * pitch + x] = inside ? inside_value : 0;
}

} // namespace ibt

IBT_API int ibt_polygon_mask(const double *poly_xy, int E, int H, int W, uint8_t *out, int64_t pitch, int inside_value,
                             void *stream)
{
    using namespace ibt;
    if (!poly_xy || !out || E < 3 || E > MAX_POLY || H <= 0 || W <= 0 || pitch < W ||
        reinterpret_cast<uintptr_t>(poly_xy) % 16 != 0)
        return IBT_E_INVALID;
    dim3 grid((W + 255) / 256, H);
    polygon_mask_kernel<<<grid, 256, (size_t)E * sizeof(double2), static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const double2 *>(poly_xy), E, H, W, out, pitch, (uint8_t)inside_value);
    return check_launch("ibt_polygon_mask");
}
