"""Generic kernel_main dispatch (rendering/_core.py:247-299).

The reference concatenates every @kernel_struct / @kernel_function / @kernel_main into one OpenCL C
program and builds it at first dispatch.  On B200 the built-in raster and ray-cast paths are hand-written
CUDA (rendertoy_b200/csrc); arbitrary user kernels are a "next" row served by an
NVRTC translation of the captured OpenCL C.  Until that lands, dispatching a generic kernel fails loudly --
there is no CPU interpreter.
"""
import math


class Dispatcher:
    def __init__(self, name, arguments, body):
        self.name, self.arguments, self.body = name, arguments, body

    def __getitem__(self, num_threads):
        if isinstance(num_threads, (list, tuple)):
            num_threads = math.prod(num_threads)

        def dispatch_call(*args):
            raise NotImplementedError(
                f"kernel_main '{self.name}': generic OpenCL-C kernels are not translated yet "
                "(built-in Raster / Raycaster paths are native CUDA; no CPU fallback exists)")

        return dispatch_call
