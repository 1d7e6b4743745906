"""Dev-time helper: print the normalised-source fingerprints of the tutorial shader pairs that
rendertoy_b200/rendering/_raster.py recognises.  Reads the reference checkout (never at run time)."""
import ast
import sys

sys.path.insert(0, ".")
from rendertoy_b200.rendering._raster import shader_fingerprint  # noqa: E402

FILES = ["tutorials/lesson08_rasterization.py", "tutorials/lesson09_texture_mapping.py",
         "Class2022/Team Camilo-Javier-Karel/scene.py"]
root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
for f in FILES:
    tree = ast.parse(open(f"{root}/{f}").read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and any("kernel_function" in ast.unparse(d) for d in node.decorator_list):
            print(f, node.name, shader_fingerprint(ast.get_docstring(node)))
