"""Compare the train-mode feature moments of the fused call (row x row block from quantize_mark) with the split call."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import helpers as H
from radardistill_b200 import _lib, ops
name = sys.argv[1] if len(sys.argv) > 1 else "dynpillar64__b2__train"
g = H.load_golden(os.path.join(H.GOLDEN_DIR, name + ".npz"))
m = H.module_from_golden(g).train()
pts = torch.from_numpy(g["points"]).cuda()
spec, pfn = m.spec, m.pfn_layers[0]
bs = int(g["points"][:, 0].max()) + 1
args = (pfn.linear.weight.detach(), None, pfn.norm.weight.detach(), pfn.norm.bias.detach(), pfn.norm.running_mean.clone(), pfn.norm.running_var.clone())
res = ops.encode_forward(pts, spec, bs, *args, True, True)
torch.cuda.synchronize()
st1 = res.bn_state.cpu().numpy().copy()
lib = _lib.load()
geom, layout = spec.geom(bs), spec.layout_struct()
prm = ops._params_struct(spec, *args, True)
feats = torch.empty((len(g["points"]), spec.c_out), device="cuda")
P = ops._ptr
_lib.check(lib.rdp_pfn_fwd(P(pts), len(g["points"]), C.byref(geom), C.byref(layout), C.byref(prm), P(res.workspace), res.workspace.numel(),
                           P(res.counters), P(feats), None, None, P(res.bn_state), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pfn")
torch.cuda.synchronize()
st2 = res.bn_state.cpu().numpy()
co = spec.c_out
G = spec.cols - 1 + (1 if spec.with_distance else 0) + 5
print("n", st1[4 * co], st2[4 * co], "G", G)
S1a, S1b = st1[4 * co + 1:4 * co + 1 + G], st2[4 * co + 1:4 * co + 1 + G]
S2a, S2b = st1[4 * co + 1 + G:4 * co + 1 + G + G * G].reshape(G, G), st2[4 * co + 1 + G:4 * co + 1 + G + G * G].reshape(G, G)
np.set_printoptions(precision=4, linewidth=200, suppress=False)
print("S1 fused", S1a); print("S1 split", S1b)
print("S2 fused - split (relative)"); print((S2a - S2b) / np.maximum(np.abs(S2b), 1e-30))
print("S2 split"); print(S2b); print("S2 fused"); print(S2a)
