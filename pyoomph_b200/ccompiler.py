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


def find_jitbridge_include() -> Optional[str]:
    """directory holding the reference's jitbridge.h: $PB2_JITBRIDGE_INCLUDE, an installed pyoomph's JIT include directory, or the
    reference tree next to this checkout; None if there is none (then the plugin exports only JIT_ELEMENT_init_cuda)"""
    cands = [os.environ.get("PB2_JITBRIDGE_INCLUDE")]
    try:
        import importlib.util
        spec = importlib.util.find_spec("pyoomph")
        if spec and spec.submodule_search_locations:
            cands += [os.path.join(p, "jitbridge") for p in spec.submodule_search_locations]
    except Exception:
        pass
    cands.append("/root/reference/src")
    for c in cands:
        if c and os.path.exists(os.path.join(c, "jitbridge.h")):
            return c
    return None


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

    # ---- pyoomph::CCompiler interface (src/ccompiler.hpp:35-76) and its Python helpers (pyoomph/generic/ccompiler.py:70-77) ------------
    code_extension = ".c"

    def set_code_from_file(self, ftrunk: str) -> None:
        """the generated source lies at <ftrunk><code_extension> (the host wrote it, src/pybind/problem.cpp:694-696)"""
        self._code_trunk = ftrunk

    def get_code_trunk(self) -> str:
        return getattr(self, "_code_trunk", "")

    def get_code_filename(self) -> str:
        return self.get_code_trunk() + self.code_extension

    def get_shared_lib_extension(self) -> str:
        return ".so"

    def get_lib_filename(self) -> str:
        return self.get_code_trunk() + self.get_shared_lib_extension()

    def get_shared_library(self, code_trunk: str) -> str:
        return code_trunk + self.get_shared_lib_extension()

    def expand_full_library_name(self, relname: str) -> str:
        return os.path.join(os.getcwd(), relname)

    def get_jit_include_dir(self) -> str:
        return INCLUDE_DIR

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
    code_extension = ".cu"

    def __init__(self):
        self.nvcc = find_nvcc()
        self.extra_flags: List[str] = []
        self.fast_math = False
        self.keep_source = True
        self.last_log = ""
        # directory of the reference's jitbridge.h (pyoomph ships it as its JIT include directory, src/ccompiler.hpp:31): when present the
        # plugin also exports the reference's own JIT_ELEMENT_init (cuda_emitter._emit_jit_element_init)
        self.jitbridge_include: Optional[str] = find_jitbridge_include()

    def check_avail(self) -> bool:
        return self.nvcc is not None

    def optimize_for_max_speed(self):
        """Counterpart of SystemCCompiler.optimize_for_max_speed (-ffast-math); off by default for parity."""
        self.fast_math = True
        return self

    def flags(self, with_jitbridge: bool = True) -> List[str]:
        fl = NVCC_ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
                          "-I", INCLUDE_DIR, "-Xptxas", "-v"]
        if self.fast_math:
            fl.append("--use_fast_math")
        else:
            fl += ["--fmad=true"]
        if self.jitbridge_include and with_jitbridge:
            fl += ["-DPB2_WITH_JITBRIDGE", "-I", self.jitbridge_include]
        return fl + self.extra_flags

    def compile_code(self, source: str, name: str, *, force: bool = False, quiet: bool = True) -> str:
        if not self.check_avail():
            raise RuntimeError("nvcc not found: the CUDA assembly path cannot be built (no CPU fallback exists)")
        os.makedirs(JIT_DIR, exist_ok=True)
        hdr = open(os.path.join(INCLUDE_DIR, "pb2_jit_cuda.h")).read()
        # the cache key leaves out where the reference's header was found: a plugin built where jitbridge.h is available (it then also
        # exports JIT_ELEMENT_init) is the same plugin on a machine without it and must not be rebuilt there
        # ... nor on where the checkout lies (the GPU box mounts the repository under another path)
        key_flags = ["<include>" if f == INCLUDE_DIR else f for f in self.flags(with_jitbridge=False)]
        tag = hashlib.sha1((source + hdr + " ".join(key_flags)).encode()).hexdigest()[:12]
        cu = os.path.join(JIT_DIR, "%s_%s.cu" % (name, tag))
        so = os.path.join(JIT_DIR, "%s_%s.so" % (name, tag))
        if os.path.exists(so) and not force:
            return so
        # several ranks may compile the same class at once: build under a private trunk, publish atomically
        tmp_trunk = os.path.join(JIT_DIR, "%s_%s.%d" % (name, tag, os.getpid()))
        with open(tmp_trunk + ".cu", "w") as f:
            f.write(source)
        self.set_code_from_file(tmp_trunk)
        try:
            self.compile(False, False, quiet, [])
        finally:
            os.replace(tmp_trunk + ".cu", cu)
            if os.path.exists(tmp_trunk + ".log"):
                os.replace(tmp_trunk + ".log", os.path.join(JIT_DIR, "%s_%s.log" % (name, tag)))
        os.replace(tmp_trunk + ".so", so)
        return so

    def compile(self, suppress_compilation: bool, suppress_code_writing: bool, quiet: bool, extra_flags: List[str]) -> bool:
        """The reference's compiler-plugin contract (src/ccompiler.hpp:69, pyoomph/generic/ccompiler.py:120-143): compile the generated
        source <trunk>.cu (set_code_from_file) into <trunk>.so with nvcc for sm_100a.  suppress_compilation: nothing is done (the host
        reuses an existing library); suppress_code_writing: the source was not rewritten, an existing <trunk>.so is kept as it is
        (like TCCBoxCompiler.compile).  Returns True; a failing nvcc raises with its log."""
        if suppress_compilation:
            return True
        if not self.check_avail():
            raise RuntimeError("nvcc not found: the CUDA assembly path cannot be built (no CPU fallback exists)")
        src, lib = self.get_code_filename(), self.get_lib_filename()
        if suppress_code_writing and os.path.exists(lib):
            return True
        if not os.path.exists(src):
            raise RuntimeError("generated CUDA source %s does not exist" % src)
        cmd = [self.nvcc] + self.flags() + list(extra_flags) + [src, "-o", lib]
        if not quiet:
            print("Compiling " + src + " with jit-include dir " + self.get_jit_include_dir())
        r = subprocess.run(cmd, capture_output=True, text=True)
        self.last_log = r.stdout + r.stderr
        with open(self.get_code_trunk() + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + self.last_log)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, self.last_log[-6000:]))
        if not quiet:
            print(self.last_log)
        return True


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
