// rt_misc.cu -- error reporting, device info and mesh upload for librendertoy_b200.so.
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
