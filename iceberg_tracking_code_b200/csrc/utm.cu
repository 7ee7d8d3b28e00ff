// K4: photo -> UTM ray/plane projection of track vertices.  Replaces the per-vertex
// Python calls cam.photocords_cropped_to_uncropped + cam.photo_to_utm at s2_cam_to_utm.py:247-254
// (imports/camtools.py:414-421, 286-332).  fp64 throughout: northing ~6.5e6 m needs it.
// The nine direction cosines are evaluated once on the host (libm, like numpy) and passed by value;
// the reference recomputes them per vertex with identical inputs, so results do not change.
#include "common.cuh"
#include <math.h>

namespace ibt {

struct UtmCam {
    double cl, ct, halfW, halfH, Hc, E0, N0;
    double sX0, sX1, sX2;                  // sigma * X
    double U0, U1, U2, V0, V1, V2;
};

__global__ void __launch_bounds__(256)
photo_to_utm_kernel(const float2 *__restrict__ xy, int64_t n, const __grid_constant__ UtmCam c, double2 *__restrict__ EN)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 p = xy[i];
        // same operation order as camtools.py:300-301, 326-330; no FMA contraction (-fmad=false)
        const double xi = ((double)p.x + c.cl) - c.halfW;
        const double yi = ((double)p.y + c.ct) - c.halfH;
        const double den = c.sX2 + xi * c.U2 + yi * c.V2;
        const double tx = c.Hc * (c.sX0 + xi * c.U0 + yi * c.V0) / den + c.E0;
        const double ty = c.Hc * (c.sX1 + xi * c.U1 + yi * c.V1) / den + c.N0;
        EN[i] = make_double2(tx, ty);
    }
}

} // namespace ibt

IBT_API int ibt_photo_to_utm(const float *xy, int64_t n, const double *cam, double *EN, void *stream)
{
    using namespace ibt;
    if (n < 0 || !cam) return IBT_E_INVALID;
    if (n == 0) return IBT_OK;
    if (!xy || !EN || reinterpret_cast<uintptr_t>(xy) % 8 != 0 || reinterpret_cast<uintptr_t>(EN) % 16 != 0) return IBT_E_INVALID;
    const double sg = cam[4], th = cam[6], ph = cam[7], ps = cam[8];
    UtmCam c;
    c.cl = cam[0]; c.ct = cam[1]; c.halfW = cam[2] / 2; c.halfH = cam[3] / 2; c.Hc = cam[5]; c.E0 = cam[9]; c.N0 = cam[10];
    c.sX0 = sg * (cos(th) * cos(ph)); c.sX1 = sg * (sin(th) * cos(ph)); c.sX2 = sg * sin(ph);
    c.U0 = sin(th) * cos(ps) - cos(th) * sin(ph) * sin(ps);
    c.U1 = -cos(th) * cos(ps) - sin(th) * sin(ph) * sin(ps);
    c.U2 = cos(ph) * sin(ps);
    c.V0 = -sin(th) * sin(ps) - cos(th) * sin(ph) * cos(ps);
    c.V1 = cos(th) * sin(ps) - sin(th) * sin(ph) * cos(ps);
    c.V2 = cos(ph) * cos(ps);
    int64_t blocks = (n + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    photo_to_utm_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2 *>(xy), n, c, reinterpret_cast<double2 *>(EN));
    return check_launch("ibt_photo_to_utm");
}
