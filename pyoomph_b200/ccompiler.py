"""Compiler plugins: the registry of pyoomph's ``BaseCCompiler`` plus the CUDA compiler that replaces tcc/gcc.

Mirrors /root/reference/pyoomph/generic/ccompiler.py: ``BaseCCompiler`` (:46) with ``register_compiler`` (:83),
``factory_compiler`` and the per-compiler ``check_avail``/``compile`` contract, ``SystemCCompiler`` (:146, gcc with
``-O3 -fPIC -march=native`` :214-220, ``-ffast-math`` via ``optimize_for_max_speed`` :233), and the C++ side
``pyoomph::CCompiler`` virtuals (src/ccompiler.hpp:35-76).  ``CudaCCompiler`` (compiler_id "cuda") turns the .cu file
the CUDA emission backend wrote into a shared object exporting ``JIT_ELEMENT_init_cuda`` with nvcc for sm_100a.
There is deliberately no CPU compiler registered here: assembly has no CPU fallback.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from typing import Dict, List, Optional, Type

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
INCLUDE_DIR = os.path.join(REPO, "include")
JIT_DIR = os.path.join(HERE, "_jit")
NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def find_nvcc() -> Optional[str]:
    for c in (os.environ.get("PB2_NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    return None


class BaseCCompiler:
    """Registry and common interface (pyoomph/generic/ccompiler.py:46-112)."""
    compiler_id: str = ""
    compiler_quality: float = 0.0
    _registry: Dict[str, Type["BaseCCompiler"]] = {}

    @classmethod
    def register_compiler(cls, *, override: bool = False):
        def deco(sub: Type["BaseCCompiler"]):
            if not sub.compiler_id:
                raise RuntimeError("compiler class needs a compiler_id")
            if sub.compiler_id in cls._registry and not override:
                raise RuntimeError("compiler id already registered: " + sub.compiler_id)
            cls._registry[sub.compiler_id] = sub
            return sub
        return deco

    @classmethod
    def factory_compiler(cls, compiler_id: str) -> "BaseCCompiler":
        if compiler_id not in cls._registry:
            raise RuntimeError("Unknown compiler id '%s' (available: %s)" % (compiler_id, sorted(cls._registry)))
        comp = cls._registry[compiler_id]()
        if not comp.check_avail():
            raise RuntimeError("Compiler '%s' is not available on this machine" % compiler_id)
        return comp

    @classmethod
    def get_available_compilers(cls) -> List[str]:
        return [k for k, v in cls._registry.items() if v().check_avail()]

    def check_avail(self) -> bool:
        return False

    def compile(self, suppress_compilation: bool, suppress_code_writing: bool, quiet: bool, extra_flags: List[str]) -> bool:
        raise NotImplementedError

    def sanity_check(self) -> bool:
        return True


@BaseCCompiler.register_compiler()
class CudaCCompiler(BaseCCompiler):
    """nvcc for sm_100a in place of tcc/gcc (src/ccompiler.cpp:71-189, pyoomph/generic/ccompiler.py:207-243).

    ``compile_code(source, name)`` is the analogue of set_code_from_file + compile + get_shared_library: it returns
    the path of the plugin .so (kept in-tree under pyoomph_b200/_jit so that it travels with the repository)."""
    compiler_id = "cuda"
    compiler_quality = 1.0

    def __init__(self):
        self.nvcc = find_nvcc()
        self.extra_flags: List[str] = []
        self.fast_math = False
        self.keep_source = True
        self.last_log = ""

    def check_avail(self) -> bool:
        return self.nvcc is not None

    def optimize_for_max_speed(self):
        """Counterpart of SystemCCompiler.optimize_for_max_speed (-ffast-math); off by default for parity."""
        self.fast_math = True
        return self

    def flags(self) -> List[str]:
        fl = NVCC_ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
                          "-I", INCLUDE_DIR, "-Xptxas", "-v"]
        if self.fast_math:
            fl.append("--use_fast_math")
        else:
            fl += ["--fmad=true"]
        return fl + self.extra_flags

    def compile_code(self, source: str, name: str, *, force: bool = False, quiet: bool = True) -> str:
        if not self.check_avail():
            raise RuntimeError("nvcc not found: the CUDA assembly path cannot be built (no CPU fallback exists)")
        os.makedirs(JIT_DIR, exist_ok=True)
        hdr = open(os.path.join(INCLUDE_DIR, "pb2_jit_cuda.h")).read()
        tag = hashlib.sha1((source + hdr + " ".join(self.flags())).encode()).hexdigest()[:12]
        cu = os.path.join(JIT_DIR, "%s_%s.cu" % (name, tag))
        so = os.path.join(JIT_DIR, "%s_%s.so" % (name, tag))
        if os.path.exists(so) and not force:
            return so
        # several ranks may compile the same class at once: build under a private name, publish atomically
        tmp_so = "%s.%d.tmp" % (so, os.getpid())
        tmp_cu = os.path.join(JIT_DIR, "%s_%s.%d.cu" % (name, tag, os.getpid()))
        with open(tmp_cu, "w") as f:
            f.write(source)
        os.replace(tmp_cu, cu)
        cmd = [self.nvcc] + self.flags() + [cu, "-o", tmp_so]
        r = subprocess.run(cmd, capture_output=True, text=True)
        self.last_log = r.stdout + r.stderr
        with open(os.path.join(JIT_DIR, "%s_%s.log" % (name, tag)), "w") as f:
            f.write(" ".join(cmd) + "\n" + self.last_log)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (cu, self.last_log[-6000:]))
        os.replace(tmp_so, so)
        if not quiet:
            print(self.last_log)
        return so

    def compile(self, suppress_compilation: bool, suppress_code_writing: bool, quiet: bool, extra_flags: List[str]) -> bool:
        raise RuntimeError("CudaCCompiler is driven through compile_code(); see INTEGRATION.md for the pyoomph-side hook")


def get_ccompiler(compiler_id: str = "cuda") -> BaseCCompiler:
    return BaseCCompiler.factory_compiler(compiler_id)


def build_core_library(force: bool = False) -> str:
    """Compile csrc/pb2_core.cu into pyoomph_b200/libpyoomph_b200.so (in-tree)."""
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found")
    src = os.path.join(HERE, "csrc", "pb2_core.cu")
    out = os.path.join(HERE, "libpyoomph_b200.so")
    deps = [src, os.path.join(INCLUDE_DIR, "pyoomph_b200.h"), os.path.join(INCLUDE_DIR, "pb2_jit_cuda.h")]
    if os.path.exists(out) and not force and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    cmd = [nvcc] + NVCC_ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fopenmp", "-shared", "-cudart", "shared",
                                "-I", INCLUDE_DIR, src, "-o", out, "-ldl", "-lgomp"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libpyoomph_b200.so failed:\n" + (r.stdout + r.stderr)[-6000:])
    return out
