"""Times rdp_index_fwd and the eval / train PFN launches on the LiDAR batch (CUDA events, L2 flushed)."""
import ctypes as C, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from radardistill_b200 import _lib, ops

def main(frames=8, reps=int(os.environ.get("RDP_BENCH_REPS", "15"))):
    dev = torch.device("cuda", 0)
    lidar, _ = bench.make_clouds(0, frames)
    lid, rad, call = bench.build_modules(dev, "B", False)
    lib = _lib.load()
    pts = torch.from_numpy(lidar).to(dev)
    spec, norm, w = lid.spec, lid.pfn_layers[0].norm, lid.pfn_layers[0].linear.weight.detach()
    geom, layout = spec.geom(frames), spec.layout_struct()
    only_eval = bool(os.environ.get('RDP_BENCH_ONLY_EVAL'))
    res = ops.encode_forward(pts, spec, frames, w, None, norm.weight, norm.bias, norm.running_mean, norm.running_var, not only_eval, not only_eval)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    feats = torch.empty((len(lidar), spec.c_out), dtype=torch.float32, device=dev)
    argp = torch.empty((len(lidar), spec.c_out), dtype=torch.int32, device=dev)
    grad = torch.ones((res.n_pillars, spec.c_out), dtype=torch.float32, device=dev)
    dw, dg, db = (torch.empty(s, device=dev) for s in ((spec.c_out, spec.c_in), (spec.c_out,), (spec.c_out,)))
    out = {}
    def timeit(name, fn):
        ts = []
        for i in range(reps + 3):
            if not os.environ.get('RDP_NO_FLUSH'): flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            if i >= 3: ts.append(a.elapsed_time(b) * 1e3)
        out[name] = round(sum(ts) / len(ts), 1)
    P = lambda t: ops._ptr(t)
    def index():
        _lib.check(lib.rdp_index_fwd(P(pts), len(lidar), C.byref(geom), spec.coord_cols, P(res.workspace), res.workspace.numel(),
                                     P(res.coords), P(res.inverse), P(res.counts), P(res.counters), st), "index")
    def pfn(train, arg):
        prm = ops._params_struct(spec, w, None, norm.weight.detach(), norm.bias.detach(), norm.running_mean.clone(), norm.running_var.clone(), train)
        keep = (prm,)
        _lib.check(lib.rdp_pfn_fwd(P(pts), len(lidar), C.byref(geom), C.byref(layout), C.byref(prm), P(res.workspace),
                                   res.workspace.numel(), P(res.counters), P(feats), P(argp) if arg else None, None,
                                   P(res.bn_state) if train else None, st), "pfn")
    def bwd():
        prm = ops._params_struct(spec, w, None, norm.weight.detach(), norm.bias.detach(), norm.running_mean, norm.running_var, True)
        _lib.check(lib.rdp_pfn_bwd(P(pts), len(lidar), C.byref(geom), C.byref(layout), C.byref(prm), P(res.workspace),
                                   res.workspace.numel(), P(res.counters), P(grad), P(feats), P(argp), P(res.bn_state),
                                   P(dw), P(dg), P(db), st), "bwd")
    timeit("index_us", index)
    timeit("pfn_eval_us", lambda: pfn(False, False))
    if not only_eval:
        timeit("pfn_train_us(stats+apply_arg)", lambda: pfn(True, True))
        timeit("pfn_bwd_us", bwd)
    out["lib"] = os.path.basename(_lib.LIB_PATH); out["N"] = res.n_kept; out["P"] = res.n_pillars
    print(json.dumps(out))

if __name__ == "__main__":
    main()
