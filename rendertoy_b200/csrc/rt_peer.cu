// rt_peer.cu -- peer-memory plumbing for the multi-GPU framebuffer gather (the reference
// is single-device, rendering/_core.py:10-11).
//
// Rank 0 owns one cudaMalloc'ed frame store and exports it with CUDA IPC; every other rank maps it and hands the
// mapped address to rt_raycast_primary / rt_raster_draw_triangles as the render target.  The shading kernels then
// store their BGRA8 pixels straight into rank 0's HBM over NVLink while they are still tracing: the "gather" is
// fused into the producing kernel and the only thing left of the collective is a barrier.
#include "rt_common.cuh"

namespace {

// One block per 32x32-pixel tile, one uint4 (four BGRA8 pixels) per thread.  The tile travels iff it holds a pixel that is
// not the clear colour, or did the last time this producer pushed into the same destination frame (then its clear pixels
// must overwrite what is there).  Reads are streaming (the frame is not needed again), the stores go straight to the
// peer-mapped frame over NVLink as whole 128-byte rows.
__global__ void __launch_bounds__(256) push_tiles_kernel(uint4 *dst, const uint4 *src, int w4, int height, uint32_t clear_px, unsigned char *state,
                                                         unsigned long long *bytes)
{
    const int r = (int)blockIdx.y * 32 + (int)(threadIdx.x >> 3), c4 = (int)blockIdx.x * 8 + (int)(threadIdx.x & 7);
    const bool in = r < height && c4 < w4;
    const size_t at = (size_t)r * (size_t)w4 + (size_t)c4;
    const unsigned tile = blockIdx.y * gridDim.x + blockIdx.x;
    const unsigned char prev = state[tile];
    uint4 v = make_uint4(clear_px, clear_px, clear_px, clear_px);
    if (in) v = __ldcs(src + at);
    const int any = __syncthreads_or(v.x != clear_px || v.y != clear_px || v.z != clear_px || v.w != clear_px);
    if (!any && !prev) return;
    if (in) __stcs(dst + at, v);
    if (threadIdx.x == 0) {
        state[tile] = any ? 1 : 0;
        if (bytes) atomicAdd(bytes, 4096ull);
    }
}

} // namespace

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

// Sparse frame push by tiles -- the gather of an animation batch whose frames are mostly content: the bounding rectangle of a
// frame-filling view is ~90 % of the frame although two thirds of its pixels are the clear colour, and at N = 8 seven
// producers pushing such rectangles saturate rank 0's NVLink ingest (~780 GB/s measured).  One kernel on the PRODUCING GPU
// reads its finished frame and stores only the 32x32-pixel tiles that hold something (or held something the last time this
// producer wrote the same destination: d_tile_state, one byte per tile, zero = "the destination tile is the clear colour")
// into the frame at d_dst (rank 0's peer-mapped slot, same layout).  The destination must start filled with clear_px
// (rt_peer_alloc zero-fills; for another colour start with a state of all ones).  d_bytes: NULL or a device uint64 the kernel
// adds the bytes it stored to.
int rt_push_tiles(void *d_dst, const void *d_src, int width, int height, uint32_t clear_px, void *d_tile_state, void *d_bytes, void *stream)
{
    RT_REQUIRE(d_dst && d_src && d_tile_state, "frames / tile state");
    RT_REQUIRE(width > 0 && height > 0 && width % 4 == 0, "frame size (width a multiple of 4 pixels)");
    RT_REQUIRE((((uintptr_t)d_dst | (uintptr_t)d_src) & 15) == 0, "16-byte aligned frames");
    dim3 grid((unsigned)((width + 31) / 32), (unsigned)((height + 31) / 32));
    push_tiles_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((uint4 *)d_dst, (const uint4 *)d_src, width / 4, height, clear_px,
                                                              (unsigned char *)d_tile_state, (unsigned long long *)d_bytes);
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

int64_t rt_push_tiles_state_bytes(int width, int height) { return (int64_t)((width + 31) / 32) * ((height + 31) / 32); }

// The image-space partition's gather: a rank that rendered its row stripes of a frame locally moves exactly
// those -- rows y in [y0, y1] with (y / stripe_rows) % mod == rem, columns [x_bytes, x_bytes + width_bytes) -- into the same
// place of the frame at d_dst (same pitch on both sides).  Whole stripes between the first and the last go as ONE 3-D
// copy-engine transfer (a stripe = a 2-D slice, slices mod * stripe_rows rows apart), the two that [y0, y1] may cut as 2-D ones.
int rt_copy_stripes(void *d_dst, const void *d_src, int64_t pitch_bytes, int64_t x_bytes, int64_t width_bytes, int64_t y0, int64_t y1,
                    int stripe_rows, int mod, int rem, void *stream)
{
    RT_REQUIRE(d_dst && d_src && pitch_bytes > 0 && x_bytes >= 0 && width_bytes >= 0 && x_bytes + width_bytes <= pitch_bytes, "frame geometry");
    RT_REQUIRE(stripe_rows >= 1 && mod >= 1 && rem >= 0 && rem < mod && y0 >= 0, "stripes");
    if (width_bytes == 0 || y1 < y0) return RT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    char *dst = (char *)d_dst + x_bytes;
    const char *src = (const char *)d_src + x_bytes;
    auto copy2d = [&](int64_t ya, int64_t yb) -> cudaError_t { // rows [ya, yb]
        return cudaMemcpy2DAsync(dst + ya * pitch_bytes, (size_t)pitch_bytes, src + ya * pitch_bytes, (size_t)pitch_bytes, (size_t)width_bytes,
                                 (size_t)(yb - ya + 1), cudaMemcpyDefault, st);
    };
    // owned stripes s = rem + j * mod that meet [y0, y1]
    const int64_t s_lo = y0 / stripe_rows, s_hi = y1 / stripe_rows;
    int64_t j0 = s_lo <= rem ? 0 : (s_lo - rem + mod - 1) / mod;
    if (rem + j0 * mod > s_hi) return RT_OK;
    const int64_t j1 = (s_hi - rem) / mod;
    int64_t ja = j0, jb = j1; // [ja, jb]: stripes wholly inside [y0, y1]
    const int64_t first = rem + j0 * (int64_t)mod, last = rem + j1 * (int64_t)mod;
    if (first * stripe_rows < y0 || (first + 1) * stripe_rows - 1 > y1) {
        const int64_t ya = first * stripe_rows > y0 ? first * stripe_rows : y0, yb = (first + 1) * stripe_rows - 1 < y1 ? (first + 1) * stripe_rows - 1 : y1;
        RT_CUDA(copy2d(ya, yb));
        ja = j0 + 1;
    }
    if (j1 >= ja && (last + 1) * stripe_rows - 1 > y1) {
        RT_CUDA(copy2d(last * stripe_rows, y1));
        jb = j1 - 1;
    }
    if (jb < ja) return RT_OK;
    const int64_t ya = (rem + ja * (int64_t)mod) * stripe_rows;
    if (jb == ja) { RT_CUDA(copy2d(ya, ya + stripe_rows - 1)); return RT_OK; }
    cudaMemcpy3DParms p = {};
    p.srcPtr = make_cudaPitchedPtr(const_cast<char *>(src) + ya * pitch_bytes, (size_t)pitch_bytes, (size_t)width_bytes, (size_t)mod * stripe_rows);
    p.dstPtr = make_cudaPitchedPtr(dst + ya * pitch_bytes, (size_t)pitch_bytes, (size_t)width_bytes, (size_t)mod * stripe_rows);
    p.extent = make_cudaExtent((size_t)width_bytes, (size_t)stripe_rows, (size_t)(jb - ja + 1));
    p.kind = cudaMemcpyDefault;
    if (cudaMemcpy3DAsync(&p, st) != cudaSuccess) { // e.g. a driver that refuses 3-D copies into IPC-mapped memory: stripe by stripe
        (void)cudaGetLastError();
        for (int64_t j = ja; j <= jb; ++j) {
            const int64_t y = (rem + j * (int64_t)mod) * stripe_rows;
            RT_CUDA(copy2d(y, y + stripe_rows - 1));
        }
    }
    return RT_OK;
}

} // extern "C"
