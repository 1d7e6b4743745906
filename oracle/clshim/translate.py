"""oracle/clshim/translate.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Turns the reference's accumulated OpenCL C program text (rendering/_core.py `__code__`) into a host shared
library: a handful of textual rewrites so it is valid C++17 over cl_compat.hpp, one launch trampoline per
__kernel, then g++ with strict float flags.  Output goes to oracle/_ref/ only (git-ignored).
"""
import ctypes
import hashlib
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(os.path.dirname(HERE), "_ref")

_VEC_TYPES = "float2|float3|float4|float8|float16|float4x4|int2|int3|int4|uint2|uint3|uint4"


def _balanced(src, i):
    """src[i] == '(' -> index just past the matching ')'."""
    depth = 0
    for j in range(i, len(src)):
        if src[j] == "(":
            depth += 1
        elif src[j] == ")":
            depth -= 1
            if depth == 0:
                return j + 1
    raise ValueError("unbalanced parentheses")


def _rewrite_int_casts(src):
    """(int)<operand> -> SatInt::cvt(<operand>): GPU-style saturating float->int (C++'s cast is undefined out of range)."""
    out, i = [], 0
    pat = re.compile(r"\(int\)\s*")
    while True:
        m = pat.search(src, i)
        if not m:
            out.append(src[i:])
            return "".join(out)
        out.append(src[i:m.start()])
        j = m.end()
        if src[j] == "(":
            k = _balanced(src, j)
        else:
            k = j
            while k < len(src) and (src[k].isalnum() or src[k] in "_.[]"):
                k += 1
            if k < len(src) and src[k] == "(":      # function call
                k = _balanced(src, k)
        out.append("SatInt::cvt(" + src[j:k] + ")")
        i = k


def to_cpp(source):
    s = source
    s = re.sub(r"#define\s+float4x4\s+float16", "", s)                      # typedef'd in cl_compat.hpp
    s = re.sub(r"\((%s)\)\s*\(" % _VEC_TYPES, lambda m: "make_%s(" % m.group(1), s)   # vector literals
    for a, b in ((".even.even", ".even_even_"), (".odd.even", ".odd_even_"), (".even.odd", ".even_odd_"), (".odd.odd", ".odd_odd_")):
        s = s.replace(a, b)
    s = _rewrite_int_casts(s)
    s = re.sub(r"\bunsigned\s+long\b", "ulong", s)
    return s


def _kernels(cpp):
    """[(name, [(ctype, is_pointer, param_name)])] for every __kernel in the program."""
    out = []
    for m in re.finditer(r"__kernel\s+void\s+(\w+)\s*\(([^)]*)\)", cpp):
        params = []
        for p in m.group(2).split(","):
            p = re.sub(r"\b(__global|__constant|__local|write_only|read_only|const)\b", " ", p).strip()
            is_ptr = "*" in p
            toks = p.replace("*", " ").split()
            params.append((" ".join(toks[:-1]), is_ptr, toks[-1]))
        out.append((m.group(1), params))
    return out


def compile_program(source):
    os.makedirs(REF_DIR, exist_ok=True)
    cpp = to_cpp(source)
    kernels = _kernels(cpp)
    tramp = []
    for name, params in kernels:
        args = []
        for i, (ctype, is_ptr, _) in enumerate(params):
            if is_ptr:
                args.append(f"({ctype}*)a[{i}]")
            elif ctype == "image2d_t":
                args.append(f"(image2d_t)a[{i}]")
            else:
                args.append(f"*({ctype}*)a[{i}]")
        tramp.append(f'extern "C" void {name}__launch(void** a, long n) {{ for (long g = 0; g < n; ++g) {{ __cl_gid = g; {name}({", ".join(args)}); }} }}')
    text = '#include "cl_compat.hpp"\n' + cpp + "\n" + "\n".join(tramp) + "\n"
    tag = hashlib.sha1(text.encode()).hexdigest()[:16]
    src_path, so_path = os.path.join(REF_DIR, f"clprog_{tag}.cpp"), os.path.join(REF_DIR, f"clprog_{tag}.so")
    if not os.path.exists(so_path):
        with open(src_path, "w") as fh:
            fh.write(text)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-w", "-Wno-psabi",
                               "-I", HERE, src_path, "-o", so_path])
    return ctypes.CDLL(so_path), {name: [(c, p) for c, p, _ in params] for name, params in kernels}
