// rt_misc.cu -- error reporting, device info and mesh upload for librendertoy_b200.so.
#include <math.h>
#include <stdarg.h>

#include "rt_common.cuh"

static thread_local char g_err[512] = "";

void rt_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int rt_sm_count()
{
    static int sm[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (sm[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        sm[dev] = n;
    }
    return sm[dev];
}

namespace {

// MeshVertex (80 B = 5 float4: P, N, C|pad, T, B; rendering/_modeling.py:22-28) -> SoA float4 P, float4 N.
// One thread per vertex, two 128-bit loads, two coalesced 128-bit stores.
__global__ void mesh_soa_kernel(const float4 *__restrict__ mesh, long long n, float4 *__restrict__ pos, float4 *__restrict__ nrm)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = __ldg(mesh + 5 * i), q = __ldg(mesh + 5 * i + 1);
    p.w = 1.0f; q.w = 0.0f;
    pos[i] = p; nrm[i] = q;
}

} // namespace

extern "C" {

int rt_abi_version(void) { return RT_ABI_VERSION; }

const char *rt_last_error(void) { return g_err; }

int rt_device_info(int *sm_count, int *l2_bytes, int *cc_major, int *cc_minor)
{
    int dev = 0;
    RT_CUDA(cudaGetDevice(&dev));
    if (sm_count) RT_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    if (l2_bytes) RT_CUDA(cudaDeviceGetAttribute(l2_bytes, cudaDevAttrL2CacheSize, dev));
    if (cc_major) RT_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_minor) RT_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    return RT_OK;
}

// Host only (no device work).  The tutorial vertex shaders place a vertex at H = (((P, 1) World) View) Proj and the
// rasterizer turns that into the pixel ((H.x / H.w + 1) W / 2, (1 - H.y / H.w) H / 2) (_raster.py:118-133); a fragment's
// pixel lies within one pixel of its triangle's screen bounding box (:239, :88-89).  So when every corner of the mesh's
// bounding box [lo, hi] is in front of the near plane (z >= 0, w > 0: no triangle is clipped), everything a draw can write
// lies in the rectangle of the eight projected corners: returned here, +- 2 px, clamped to the frame (possibly empty:
// x1 < x0).  Returns 1, or 0 when there is no such bound (box reaches the near plane, non-finite data).
// What it is for: a frame is the clear colour outside the union of its draws' rectangles, so only that part has to
// be read back or gathered (Raster.content_rect).
int rt_raster_screen_bounds(const float *globals48, const double *lo, const double *hi, int width, int height, int *rect)
{
    if (!globals48 || !lo || !hi || !rect || width <= 0 || height <= 0) return 0;
    double smin = INFINITY, smax = -INFINITY, tmin = INFINITY, tmax = -INFINITY, wmax = 0.0, wmin = INFINITY;
    for (int k = 0; k < 8; ++k) {
        double v[4] = {(k & 4) ? hi[0] : lo[0], (k & 2) ? hi[1] : lo[1], (k & 1) ? hi[2] : lo[2], 1.0};
        for (int m = 0; m < 3; ++m) {
            const float *M = globals48 + 16 * m;
            double r[4];
            for (int j = 0; j < 4; ++j) r[j] = v[0] * M[j] + v[1] * M[4 + j] + v[2] * M[8 + j] + v[3] * M[12 + j];
            v[0] = r[0]; v[1] = r[1]; v[2] = r[2]; v[3] = r[3];
        }
        if (!(v[0] - v[0] == 0.0) || !(v[1] - v[1] == 0.0) || !(v[2] - v[2] == 0.0) || !(v[3] - v[3] == 0.0)) return 0;
        if (!(v[2] > 0.0) || !(v[3] > 0.0)) return 0; // at or behind the near plane: clipping, no bound
        const double s = v[0] / v[3], t = v[1] / v[3];
        smin = s < smin ? s : smin; smax = s > smax ? s : smax; tmin = t < tmin ? t : tmin; tmax = t > tmax ? t : tmax;
        wmax = v[3] > wmax ? v[3] : wmax; wmin = v[3] < wmin ? v[3] : wmin;
    }
    if (!(wmin > 1e-6 * wmax)) return 0; // float32 vertex arithmetic would not resolve this
    const double big = 1073741824.0;
    auto clampd = [&](double x) { return x < -big ? -big : (x > big ? big : x); };
    const double px0 = clampd(floor((smin + 1.0) * (width * 0.5)) - 2.0), px1 = clampd(floor((smax + 1.0) * (width * 0.5)) + 3.0);
    const double py0 = clampd(floor((1.0 - tmax) * (height * 0.5)) - 2.0), py1 = clampd(floor((1.0 - tmin) * (height * 0.5)) + 3.0);
    rect[0] = px0 > 0.0 ? (int)px0 : 0; rect[1] = py0 > 0.0 ? (int)py0 : 0;
    rect[2] = px1 < width - 1 ? (int)px1 : width - 1; rect[3] = py1 < height - 1 ? (int)py1 : height - 1;
    return 1;
}

// The same for n boxes (lo / hi: n x 3 doubles): the union of their rectangles.  A mesh's own bounding box projects to a
// rectangle a quarter larger than the union of the rectangles of 64 chunks of it (dragon orbit: 0.228 -> 0.175 of the 4K frame
// for the lesson06 camera, 0.748 -> 0.594 for lesson08), and that rectangle is what every sparse read-back and gather moves.
// Returns 1 (x1 < x0: nothing on screen), or 0 as soon as one box has no bound.
int rt_raster_screen_bounds_n(const float *globals48, const double *lo, const double *hi, int n_boxes, int width, int height, int *rect)
{
    if (!rect || n_boxes < 1) return 0;
    int u[4] = {width, height, -1, -1}, r[4];
    for (int i = 0; i < n_boxes; ++i) {
        if (!rt_raster_screen_bounds(globals48, lo + 3 * i, hi + 3 * i, width, height, r)) return 0;
        if (r[2] < r[0] || r[3] < r[1]) continue;
        u[0] = r[0] < u[0] ? r[0] : u[0]; u[1] = r[1] < u[1] ? r[1] : u[1];
        u[2] = r[2] > u[2] ? r[2] : u[2]; u[3] = r[3] > u[3] ? r[3] : u[3];
    }
    for (int k = 0; k < 4; ++k) rect[k] = u[k];
    return 1;
}

int rt_mesh_upload_soa(const void *d_mesh_vertices, int64_t n_vertices, void *d_pos4, void *d_nrm4, void *stream)
{
    RT_REQUIRE(n_vertices >= 0, "vertex count");
    if (n_vertices == 0) return RT_OK;
    RT_REQUIRE(d_mesh_vertices && d_pos4 && d_nrm4, "buffers");
    RT_REQUIRE((((uintptr_t)d_mesh_vertices | (uintptr_t)d_pos4 | (uintptr_t)d_nrm4) & 15) == 0, "16-byte alignment");
    mesh_soa_kernel<<<(unsigned)((n_vertices + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4 *)d_mesh_vertices, n_vertices,
                                                                                           (float4 *)d_pos4, (float4 *)d_nrm4);
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

} // extern "C"
