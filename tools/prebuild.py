#!/usr/bin/env python
"""Pre-compile plugin variants here (nvcc cross-compiles) so a GPU-box sweep does not spend box time in nvcc.
usage: python tools/prebuild.py [--workload ns_cavity] "ENV1=a ENV2=b" "ENV1=c" ...   (one variant per argument)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
wl = "ns_cavity"
if args and args[0] == "--workload":
    wl = args[1]; args = args[2:]
SNIP = """
import sys, re; sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
import bench
from pyoomph_b200.ccompiler import get_ccompiler
from pyoomph_b200.cuda_emitter import CudaEmitter
pb = bench.build_workload(%r, 4)
code = pb['code']
cc = get_ccompiler('cuda')
so = cc.compile_code(CudaEmitter(code, code.name).emit(), code.name)
log = open(so[:-3] + '.log').read()
m = re.search(r"Compiling entry function 'pb2_\\w+_r0_f1'.*?Used (\\d+) registers.*?\\n", log, re.S)
sp = re.search(r"Compiling entry function 'pb2_\\w+_r0_f1'.*?(\\d+) bytes spill stores", log, re.S)
print(so.split('/')[-1], 'regs', m.group(1) if m else '?', 'spill', sp.group(1) if sp else '?')
""" % (ROOT, ROOT, wl)
procs = []
for v in args:
    env = dict(os.environ)
    for kv in v.split():
        k, _, val = kv.partition("=")
        env[k] = val
    procs.append((v, subprocess.Popen([sys.executable, "-c", SNIP], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for v, p in procs:
    out = p.communicate()[0]
    print("[%s] %s" % (v, out.strip().splitlines()[-1] if out.strip() else "no output"))
