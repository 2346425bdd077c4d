"""The reference's plugin contract, exercised from C exactly like the reference's host does (tests/c/plugin_contract.c, compiled against
/root/reference/src/jitbridge.h): zero-filled table, check_compiler_size handshake, JIT_ELEMENT_init by dlsym, metadata against
JIT_ELEMENT_init_cuda's pb2_class_info, clean_up; and the CudaCCompiler.compile(...) contract.  CPU part: no device needed (JIT_ELEMENT_init
is host code).  GPU part: the same program also runs one assembly through pb2_problem_assemble_host and must reproduce B200Assembly's
numbers bit for bit."""
import os
import shutil
import struct
import subprocess
import sys

import numpy as np
import pytest

from problems import make_gpu, make_problem

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_SRC = "/root/reference/src"
HAVE_REF = os.path.exists(os.path.join(REF_SRC, "jitbridge.h"))
BIN = os.path.join(HERE, "c", "plugin_contract")


def _build_program():
    """the checked-in program is compiled where the reference's header is available; the binary (git-ignored) travels to the GPU box"""
    src = os.path.join(HERE, "c", "plugin_contract.c")
    deps = [src, os.path.join(ROOT, "include", "pb2_jit_cuda.h"), os.path.join(ROOT, "include", "pyoomph_b200.h")]
    if HAVE_REF and (not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(d) for d in deps)):
        subprocess.run(["gcc", "-O1", "-std=gnu99", "-I", REF_SRC, "-I", os.path.join(ROOT, "include"), src, "-o", BIN, "-ldl"], check=True)
    return os.path.exists(BIN)


def _plugin(kind):
    from pyoomph_b200.ccompiler import get_ccompiler
    from pyoomph_b200.cuda_emitter import CudaEmitter
    pb = make_problem(kind, 3)
    cc = get_ccompiler("cuda")
    return pb, cc.compile_code(CudaEmitter(pb["code"], pb["code"].name).emit(), pb["code"].name)


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference's jitbridge.h")
@pytest.mark.parametrize("kind", ["ns_param", "ale_axi_obs", "heat3d", "ns_axi_swirl"])
def test_reference_loader_accepts_the_cuda_plugin(kind):
    assert _build_program()
    pb, so = _plugin(kind)
    sym = subprocess.run(["nm", "-D", so], capture_output=True, text=True).stdout
    assert " T JIT_ELEMENT_init\n" in sym and " T JIT_ELEMENT_init_cuda\n" in sym
    r = subprocess.run([BIN, so, os.path.join(ROOT, "pyoomph_b200", "libpyoomph_b200.so")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "contract ok" in r.stdout


def test_cuda_ccompiler_compile_contract(tmp_path):
    """compile(suppress_compilation, suppress_code_writing, quiet, extra_flags) on <trunk>.cu (src/ccompiler.hpp:69,
    pyoomph/generic/ccompiler.py:120-143)"""
    from pyoomph_b200.ccompiler import BaseCCompiler, get_ccompiler
    from pyoomph_b200.cuda_emitter import CudaEmitter
    cc = get_ccompiler("cuda")
    assert isinstance(cc, BaseCCompiler) and cc.compiler_id == "cuda" and "cuda" in BaseCCompiler.get_available_compilers()
    pb = make_problem("poisson", 2)
    trunk = str(tmp_path / "poisson_code")
    cc.set_code_from_file(trunk)
    assert cc.get_code_filename() == trunk + ".cu" and cc.get_lib_filename() == trunk + ".so" == cc.get_shared_library(trunk)
    assert cc.compile(True, False, True, []) is True and not os.path.exists(trunk + ".so")        # suppressed: nothing happens
    with pytest.raises(RuntimeError):
        cc.compile(False, False, True, [])                                                         # no source written yet
    with open(trunk + ".cu", "w") as f:
        f.write(CudaEmitter(pb["code"], "poisson").emit())
    assert cc.compile(False, False, True, ["-DPB2_EXTRA_FLAG_SEEN"]) is True and os.path.exists(trunk + ".so")
    assert "-DPB2_EXTRA_FLAG_SEEN" in open(trunk + ".log").read()
    t0 = os.path.getmtime(trunk + ".so")
    assert cc.compile(False, True, True, []) is True and os.path.getmtime(trunk + ".so") == t0     # code not rewritten: library kept
    with open(trunk + ".cu", "w") as f:
        f.write("this is not CUDA\n")
    with pytest.raises(RuntimeError):
        cc.compile(False, False, True, [])


@pytest.mark.gpu
def test_c_host_assembles_through_the_c_abi(tmp_path):
    """the C program loads plugin and engine with dlopen like a host application and assembles: bit-identical with B200Assembly"""
    if not os.path.exists(BIN) and not _build_program():
        pytest.skip("plugin_contract binary not built (needs the reference's jitbridge.h at build time)")
    pb, so = _plugin("ns")
    mesh, dm = pb["mesh"], pb["dofmap"]
    blob, out = str(tmp_path / "problem.bin"), str(tmp_path / "out.bin")
    with open(blob, "wb") as f:
        f.write(struct.pack("4q", mesh.n_elem, mesh.n_node, dm.n_dof, mesh.elem_nodes.shape[1]))
        f.write(np.ascontiguousarray(mesh.elem_nodes, dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(dm.node_eqn, dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(mesh.node_pos, dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(pb["vals"][0], dtype=np.float64).tobytes())
    r = subprocess.run([BIN, so, os.path.join(ROOT, "pyoomph_b200", "libpyoomph_b200.so"), blob, out], capture_output=True, text=True)
    assert r.returncode == 0 and "assembly ok" in r.stdout, r.stdout + r.stderr
    asm = make_gpu(pb)
    asm.assemble(flag=1)
    res, jac, _ = asm.fetch()
    raw = open(out, "rb").read()
    nnz = struct.unpack("q", raw[:8])[0]
    assert nnz == asm.nnz
    o = 8
    rs = np.frombuffer(raw, dtype=np.int32, count=dm.n_dof + 1, offset=o); o += 4 * (dm.n_dof + 1)
    ci = np.frombuffer(raw, dtype=np.int32, count=nnz, offset=o); o += 4 * nnz
    r_c = np.frombuffer(raw, dtype=np.float64, count=dm.n_dof, offset=o); o += 8 * dm.n_dof
    j_c = np.frombuffer(raw, dtype=np.float64, count=nnz, offset=o)
    assert np.array_equal(rs, asm.indptr) and np.array_equal(ci, asm.indices)
    assert np.array_equal(r_c, res) and np.array_equal(j_c, jac)
    asm.close()


def test_prebuilt_plugins_are_reused_where_the_reference_header_is_absent(tmp_path):
    """the GPU box has no /root/reference: the cache key of a plugin must not depend on where jitbridge.h was found at build time, or every
    class would be recompiled there (and lose JIT_ELEMENT_init).  A compiler without the header and WITHOUT a working nvcc must hand back
    the very same prebuilt shared object."""
    from pyoomph_b200.ccompiler import CudaCCompiler
    from pyoomph_b200.cuda_emitter import CudaEmitter
    pb, so = _plugin("poisson")
    cc = CudaCCompiler()
    cc.jitbridge_include = None
    cc.nvcc = "/nonexistent/nvcc"
    assert cc.compile_code(CudaEmitter(pb["code"], pb["code"].name).emit(), pb["code"].name) == so
    # ... and the key must not contain the absolute path of the checkout either (the box mounts it elsewhere)
    import pyoomph_b200.ccompiler as ccm
    saved = ccm.INCLUDE_DIR
    try:
        ccm.INCLUDE_DIR = str(tmp_path / "include")
        shutil.copytree(saved, ccm.INCLUDE_DIR)
        assert os.path.basename(cc.compile_code(CudaEmitter(pb["code"], pb["code"].name).emit(), pb["code"].name)) == os.path.basename(so)
    finally:
        ccm.INCLUDE_DIR = saved
