#!/usr/bin/env python
"""GPU diagnostic: capture one UNet forward of the plan into a CUDA graph (torch.cuda.CUDAGraph(keep_graph=True)) and dump
it as a dot file; prints how many kernel->kernel edges are programmatic (PDL) edges."""
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cv_diffusion_model_b200.engine import Engine  # noqa: E402
from tests.util import seeded_unet  # noqa: E402

os.makedirs("gpurun_out", exist_ok=True)
m = seeded_unet("small", 256)
eng = Engine(m, 2, 64, 64, precision="bf16", device="cuda")
x = torch.randn(2, 6, 64, 64, device="cuda")
t = torch.full((2,), 499, device="cuda", dtype=torch.long)
for _ in range(2):
    eng.forward(x, t)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph(keep_graph=True)
with torch.cuda.graph(g):
    y = eng.forward(x, t)
path = "gpurun_out/forward_graph.dot"
g.debug_dump(path)
txt = open(path).read() if os.path.exists(path) else ""
edges = re.findall(r"->", txt)
print("dot bytes", len(txt), "edges", len(edges), "lines mentioning programmatic:", len(re.findall(r"(?i)programmatic", txt)))
for line in txt.splitlines():
    if "->" in line:
        print("sample edge:", line.strip()[:200])
        break
for line in txt.splitlines():
    if re.search(r"(?i)programmatic", line):
        print("sample programmatic:", line.strip()[:200])
        break

# edge types straight from the runtime: cudaGraphEdgeData { u8 from_port, to_port, type, reserved[5] }; type 1 = programmatic
import ctypes as C
import glob
cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + glob.glob("/usr/local/cuda/lib64/libcudart.so*")
rt = C.CDLL(cands[0])
graph = C.c_void_p(g.raw_cuda_graph())
n = C.c_size_t(0)
fn = getattr(rt, "cudaGraphGetEdges_v2")
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]
rc = fn(graph, None, None, None, C.byref(n))
print("cudaGraphGetEdges_v2 rc", rc, "edges", n.value)
fr = (C.c_void_p * n.value)(); to = (C.c_void_p * n.value)(); ed = (C.c_uint8 * (8 * n.value))()
rc = fn(graph, fr, to, ed, C.byref(n))
types = [ed[8 * i + 2] for i in range(n.value)]
ports = [ed[8 * i] for i in range(n.value)]
print("rc", rc, "edge types:", {t: types.count(t) for t in set(types)}, "from_ports:", {t: ports.count(t) for t in set(ports)})
