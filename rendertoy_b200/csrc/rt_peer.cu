// rt_peer.cu -- peer-memory plumbing for the multi-GPU framebuffer gather (the reference
// is single-device, rendering/_core.py:10-11).
//
// Rank 0 owns one cudaMalloc'ed frame store and exports it with CUDA IPC; every other rank maps it and hands the
// mapped address to rt_raycast_primary / rt_raster_draw_triangles as the render target.  The shading kernels then
// store their BGRA8 pixels straight into rank 0's HBM over NVLink while they are still tracing: the "gather" is
// fused into the producing kernel and the only thing left of the collective is a barrier.
#include "rt_common.cuh"

extern "C" {

int rt_peer_alloc(int64_t bytes, void **out_d_ptr)
{
    RT_REQUIRE(bytes > 0 && out_d_ptr, "size / out pointer");
    RT_CUDA(cudaMalloc(out_d_ptr, (size_t)bytes));
    RT_CUDA(cudaMemset(*out_d_ptr, 0, (size_t)bytes)); // a cleared frame store: rt_copy_rect pushes only what can differ from it
    RT_CUDA(cudaDeviceSynchronize());                  // ... and cleared before the handle leaves this process
    return RT_OK;
}

int rt_peer_free(void *d_ptr)
{
    if (d_ptr) RT_CUDA(cudaFree(d_ptr));
    return RT_OK;
}

int rt_peer_export(const void *d_ptr, void *handle64)
{
    RT_REQUIRE(d_ptr && handle64, "pointer / handle buffer (64 bytes)");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    RT_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
    memcpy(handle64, &h, 64);
    return RT_OK;
}

int rt_peer_open(const void *handle64, void **out_d_ptr)
{
    RT_REQUIRE(handle64 && out_d_ptr, "handle / out pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    RT_CUDA(cudaIpcOpenMemHandle(out_d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return RT_OK;
}

int rt_peer_close(void *d_ptr)
{
    if (d_ptr) RT_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return RT_OK;
}

// Sparse frame movement: a ray-cast frame differs from the clear colour only inside the scene's projected bounds
// (Raycaster.screen_bounds, a quarter of the dragon's 4K frame), so only that pixel rectangle -- `rows` rows of
// `width_bytes`, row pitches in bytes -- has to travel: from a rank's local frame into its slot of rank 0's frame store
// (peer-mapped, over NVLink) while its SMs trace the next frame, or into a pinned host frame (the read-back; pyopencl's
// enqueue_copy(queue, array, image, origin, region) on the reference side).  ONE pitched copy-engine transfer on `stream`;
// either side may be local device, peer-mapped device or pinned host memory (cudaMemcpyDefault).
int rt_copy_rect(void *d_dst, int64_t dst_pitch_bytes, const void *d_src, int64_t src_pitch_bytes, int64_t width_bytes, int64_t rows,
                      void *stream)
{
    RT_REQUIRE(width_bytes >= 0 && rows >= 0, "rectangle size");
    if (width_bytes == 0 || rows == 0) return RT_OK;
    RT_REQUIRE(d_dst && d_src, "source / destination");
    RT_REQUIRE(dst_pitch_bytes >= width_bytes && src_pitch_bytes >= width_bytes, "row pitch smaller than the row");
    RT_CUDA(cudaMemcpy2DAsync(d_dst, (size_t)dst_pitch_bytes, d_src, (size_t)src_pitch_bytes, (size_t)width_bytes, (size_t)rows,
                              cudaMemcpyDefault, (cudaStream_t)stream));
    return RT_OK;
}

} // extern "C"
