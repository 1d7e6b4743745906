"""oracle/clshim/build_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Recipe run by __graft_entry__.build() when /root/reference is present: imports the unmodified reference package
through fake_cl, instantiates its Raster with the lesson08 and lesson09 tutorial shaders (one process each, as
the reference keeps one global OpenCL program per process) and compiles the resulting OpenCL C program text
for the host CPU.  Outputs: oracle/_ref/clprog_*.{cpp,so} (git-ignored).  Reference sources are read where they
lie; nothing is copied into the repo.  tests/golden/*.npz are produced by make_golden.py from these programs.

It also byte-compiles the nine tutorial scripts (tutorials/lesson01 .. lesson09) where they lie into
oracle/_ref/tutorials/*.pycode (a .pyc under another extension: the gpurun snapshot drops *.pyc) -- the Python counterpart of compiling a C reference into oracle/_ref/*.so: the GPU box has
no /root/reference, and tests/test_tutorials_gpu.py executes the UNMODIFIED tutorial programs against this package
(runpy on the .py here, exec of the unmarshalled code object there).  They are git-ignored build outputs like the .so files.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("RENDERTOY_REFERENCE", "/root/reference")

CHILD = r'''
import sys
sys.path.insert(0, {repo!r})
sys.argv = ["make_golden"]
import importlib.util
spec = importlib.util.spec_from_file_location("mg", {mg!r})
mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
ren = mg.ren
lesson = {lesson}
ns = mg.tutorial_definitions({ref!r} + "/tutorials/lesson%02d_" % lesson + ("rasterization.py" if lesson == 8 else "texture_mapping.py"))
target = ren.create_image2d(64, 48, ren._core.RGBA)
g = ren.create_struct(ns["Transforms"])
fg = g if lesson == 8 else ren.create_struct(ns["Materials"])
ren.Raster(target, ns["transform_and_draw"], g, ns["fragment_to_color"], fg)
import pyopencl as cl
prog = cl.Program(ren._core.__ctx__, ren._core.__code__).build()
print("lesson%02d: %d kernels compiled from the reference's program text" % (lesson, len(prog.kernels)))
'''


TUTORIALS = ("lesson01_math", "lesson02_vectors_and_matrices", "lesson03_drawing_images", "lesson04_mandelbrot_animation",
             "lesson05_drawing_points", "lesson06_loading_obj", "lesson07_generative_modeling", "lesson08_rasterization",
             "lesson09_texture_mapping")


def compile_tutorials():
    import py_compile
    out = os.path.join(REPO, "oracle", "_ref", "tutorials")
    os.makedirs(out, exist_ok=True)
    for name in TUTORIALS:
        src = os.path.join(REFERENCE, "tutorials", name + ".py")
        if os.path.exists(src):
            py_compile.compile(src, cfile=os.path.join(out, name + ".pycode"), dfile=f"<reference>/tutorials/{name}.py", doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    print("tutorials: byte-compiled", len(TUTORIALS), "scripts -> oracle/_ref/tutorials/")


def main():
    if not os.path.isdir(REFERENCE):
        print("reference checkout not present; nothing to build")
        return 0
    for f in glob.glob(os.path.join(REPO, "oracle", "_ref", "clprog_*")):
        os.remove(f)
    compile_tutorials()
    for lesson in (8, 9):
        code = CHILD.format(repo=REPO, mg=os.path.join(HERE, "make_golden.py"), ref=REFERENCE, lesson=lesson)
        subprocess.check_call([sys.executable, "-c", code], cwd="/tmp")
    return 0


if __name__ == "__main__":
    sys.exit(main())
