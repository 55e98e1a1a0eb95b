// rdp_torch.cpp -- native host path of the paired pillar-encoding step (torch C++ extension, built in-tree by
// radardistill_b200/build.py as rdp_torch_ext*.so).
//
// One call enqueues BOTH encoders of PillarNet.forward (pcdet/models/detectors/pillarnet.py:28-33: `vfe` then `radar_vfe`)
// on two streams, waits for the early (N, P) publication of each, narrows the outputs and wires ONE autograd node whose
// backward launches both rdp_pfn_bwd calls -- the work radardistill_b200/ops.py + vfe.forward_pair do in ~480 us of Python
// per step (measured, tools/dbg_scale.py), here in C++.  Plumbing only: device memory, streams, events, autograd; every
// computation is a librdp.so kernel (include/rdp.h).  Results are identical to the Python path (same library calls).
#include <cuda_runtime.h>
#include <torch/extension.h>

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDACachingAllocator.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <mutex>
#include <vector>

#include "rdp.h"

namespace {

using torch::Tensor;

void check(int status, const char *what) {
    if (status != RDP_OK) {
        std::string msg = std::string(what) + " failed: " + rdp_status_string(status);
        if (status == RDP_ERR_CUDA) msg += std::string(" [") + rdp_last_cuda_error() + "]";
        throw std::runtime_error(msg);
    }
}

// ---- pinned (N, P) mailboxes + events, pooled per device
struct Mailbox {
    int32_t *host = nullptr;
    cudaEvent_t event = nullptr;
};
std::mutex g_pool_mutex;
std::vector<std::vector<Mailbox>> g_pool(64);

Mailbox take_mailbox(int dev) {
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        auto &p = g_pool[dev];
        if (!p.empty()) {
            Mailbox m = p.back();
            p.pop_back();
            return m;
        }
    }
    Mailbox m;
    C10_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void **>(&m.host), sizeof(int32_t) * RDP_NUM_COUNTERS, cudaHostAllocPortable | cudaHostAllocMapped));
    C10_CUDA_CHECK(cudaEventCreateWithFlags(&m.event, cudaEventDisableTiming));
    return m;
}
void give_mailbox(int dev, Mailbox m) {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    g_pool[dev].push_back(m);
}

// ---- one encoder of the pair, as handed over by the Python module (radardistill_b200/vfe.py)
struct Enc {
    rdp_geom_t geom;
    rdp_layout_t layout;
    double eps, momentum;
    // layout of the call's single scratch allocation (ops._Plan): [librdp workspace | counters | bn_state | inverse | counts]
    int64_t ws_bytes, off_counters, off_bn, off_inverse, off_counts, total_bytes, cap;
    Tensor points, offsets;                                   // offsets: undefined => rows carry the batch column
    Tensor weight, bias, gamma, beta, rmean, rvar, nbt;       // undefined where the module has none
    bool train_bn, want_grad;
};

Enc unpack(const py::tuple &t) {
    Enc e;
    auto geom = t[0].cast<std::vector<double>>();     // lo(3) vsz(3) off(3) nx ny batch cols nz
    auto lay = t[1].cast<std::vector<int64_t>>();     // layout use_abs use_cluster use_relative with_distance c_in c_out coord_cols
    auto plan = t[2].cast<std::vector<int64_t>>();    // ws_bytes off_counters off_bn off_inverse off_counts total_bytes cap
    for (int k = 0; k < 3; ++k) { e.geom.lo[k] = (float)geom[k]; e.geom.vsz[k] = (float)geom[3 + k]; e.geom.off[k] = (float)geom[6 + k]; }
    e.geom.nx = (int32_t)geom[9]; e.geom.ny = (int32_t)geom[10]; e.geom.batch_size = (int32_t)geom[11];
    e.geom.cols = (int32_t)geom[12]; e.geom.nz = (int32_t)geom[13];
    e.layout.layout = (int32_t)lay[0]; e.layout.use_abs = (int32_t)lay[1]; e.layout.use_cluster = (int32_t)lay[2];
    e.layout.use_relative = (int32_t)lay[3]; e.layout.with_distance = (int32_t)lay[4]; e.layout.c_in = (int32_t)lay[5];
    e.layout.c_out = (int32_t)lay[6]; e.layout.coord_cols = (int32_t)lay[7];
    e.ws_bytes = plan[0]; e.off_counters = plan[1]; e.off_bn = plan[2]; e.off_inverse = plan[3]; e.off_counts = plan[4];
    e.total_bytes = plan[5]; e.cap = plan[6];
    e.eps = t[3].cast<double>(); e.momentum = t[4].cast<double>();
    auto opt = [&](int i) { return t[i].is_none() ? Tensor() : t[i].cast<Tensor>(); };
    e.points = t[5].cast<Tensor>(); e.offsets = opt(6);
    e.weight = t[7].cast<Tensor>(); e.bias = opt(8); e.gamma = opt(9); e.beta = opt(10); e.rmean = opt(11); e.rvar = opt(12); e.nbt = opt(13);
    e.train_bn = t[14].cast<bool>(); e.want_grad = t[15].cast<bool>();
    return e;
}

void check_param(const Tensor &p, std::initializer_list<int64_t> shape, const char *name, const c10::Device &dev) {
    if (!p.defined()) return;
    TORCH_CHECK(p.is_cuda() && p.device() == dev && p.scalar_type() == torch::kFloat32 && p.is_contiguous() && p.sizes() == c10::IntArrayRef(shape),
                name, ": expected a contiguous CUDA float32 tensor of the module's shape on ", dev);
}

rdp_pfn_params_t params_of(const Enc &e, bool train_bn) {
    rdp_pfn_params_t p;
    memset(&p, 0, sizeof(p));
    auto f = [](const Tensor &t) -> float * { return t.defined() ? t.data_ptr<float>() : nullptr; };
    p.weight = f(e.weight); p.bias = f(e.bias); p.gamma = f(e.gamma); p.beta = f(e.beta);
    p.running_mean = f(e.rmean); p.running_var = f(e.rvar);
    p.eps = e.eps; p.momentum = e.momentum; p.train_bn = train_bn ? 1 : 0;
    p.num_batches_tracked = (train_bn && e.nbt.defined()) ? e.nbt.data_ptr<int64_t>() : nullptr;
    return p;
}

struct Launched {
    Tensor buf, coords, features, argpos, points;
    Mailbox box;
    bool train_bn;
};

// enqueue index + PFN forward of one encoder on the CURRENT stream (ops.encode_launch)
Launched launch(Enc &e) {
    TORCH_CHECK(e.points.is_cuda(), "the pillar encoder has no CPU path: `points` must be a CUDA tensor");
    const int in_cols = e.geom.cols - (e.offsets.defined() ? 1 : 0);
    TORCH_CHECK(e.points.dim() == 2 && e.points.size(1) == in_cols, "points must be (N, ", in_cols, ")");
    Tensor pts = e.points.detach();
    if (pts.scalar_type() != torch::kFloat32) pts = pts.to(torch::kFloat32);
    if (!pts.is_contiguous()) pts = pts.contiguous();
    if (reinterpret_cast<uintptr_t>(pts.data_ptr()) % 16) pts = pts.clone();
    const auto dev = pts.device();
    const bool use_norm = e.gamma.defined();
    const int64_t co = e.layout.c_out, ci = e.layout.c_in;
    check_param(e.weight, {co, ci}, "linear.weight", dev); check_param(e.bias, {co}, "linear.bias", dev);
    check_param(e.gamma, {co}, "norm.weight", dev); check_param(e.beta, {co}, "norm.bias", dev);
    check_param(e.rmean, {co}, "norm.running_mean", dev); check_param(e.rvar, {co}, "norm.running_var", dev);
    if (e.offsets.defined())
        TORCH_CHECK(e.offsets.is_cuda() && e.offsets.scalar_type() == torch::kInt32 && e.offsets.dim() == 1 && e.offsets.is_contiguous() &&
                        e.offsets.size(0) == e.geom.batch_size + 1, "frame offsets must be a contiguous CUDA int32 tensor with batch_size + 1 entries");
    Launched L;
    L.train_bn = e.train_bn && use_norm;
    const int64_t n0 = pts.size(0);
    auto u8 = torch::TensorOptions().dtype(torch::kUInt8).device(dev);
    L.buf = torch::empty({e.total_bytes}, u8);
    L.coords = torch::empty({e.cap, e.layout.coord_cols}, u8.dtype(torch::kInt32));
    L.features = torch::empty({e.cap, co}, u8.dtype(torch::kFloat32));
    if (e.want_grad) L.argpos = torch::empty({e.cap, co}, u8.dtype(torch::kInt32));
    L.points = pts;
    L.box = take_mailbox(dev.index());
    auto st = at::cuda::getCurrentCUDAStream(dev.index()).stream();
    rdp_pfn_params_t prm = params_of(e, L.train_bn);
    char *base = static_cast<char *>(L.buf.data_ptr());
    check(rdp_encode_fwd_frames(pts.data_ptr<float>(), e.offsets.defined() ? e.offsets.data_ptr<int32_t>() : nullptr, n0, &e.geom, &e.layout,
                                &prm, base, (size_t)e.ws_bytes, L.coords.data_ptr<int32_t>(), reinterpret_cast<int32_t *>(base + e.off_inverse),
                                reinterpret_cast<int32_t *>(base + e.off_counts), reinterpret_cast<int32_t *>(base + e.off_counters),
                                L.features.data_ptr<float>(), L.argpos.defined() ? L.argpos.data_ptr<int32_t>() : nullptr,
                                L.train_bn ? reinterpret_cast<double *>(base + e.off_bn) : nullptr, L.box.host, L.box.event, st),
          "rdp_encode_fwd_frames");
    return L;
}

// wait for the early (N, P) publication and narrow (ops.encode_finish)
struct Finished {
    Tensor features, coords, argpos;
    int64_t n_kept, n_pillars;
};
Finished finish(const Enc &e, Launched &L) {
    C10_CUDA_CHECK(cudaEventSynchronize(L.box.event));
    volatile int32_t *h = L.box.host;
    Finished f;
    f.n_kept = h[RDP_CNT_N]; f.n_pillars = h[RDP_CNT_P];
    const int err = h[RDP_CNT_ERRFLAGS];
    give_mailbox(L.points.device().index(), L.box);
    TORCH_CHECK_VALUE(!(err & 1), "points[:, 0] holds a batch index outside [0, ", e.geom.batch_size, ")");
    f.features = L.features.narrow(0, 0, f.n_pillars);
    f.coords = L.coords.narrow(0, 0, f.n_pillars);
    if (L.argpos.defined()) f.argpos = L.argpos.narrow(0, 0, f.n_pillars);
    return f;
}

// parameter gradients of one encoder on the CURRENT stream (ops.encode_backward); returns [dW, dgamma | undefined, dbeta_or_dbias]
std::vector<Tensor> backward_one(const rdp_geom_t &geom, const rdp_layout_t &layout, const rdp_pfn_params_t &prm, const Tensor &points,
                                 const Tensor &buf, int64_t ws_bytes, int64_t off_counters, int64_t off_bn, bool train_bn, const Tensor &argpos,
                                 Tensor g) {
    const auto dev = points.device();
    if (g.scalar_type() != torch::kFloat32) g = g.to(torch::kFloat32);
    if (!g.is_contiguous()) g = g.contiguous();
    if (reinterpret_cast<uintptr_t>(g.data_ptr()) % 16) g = g.clone();   // the backward stages gradient rows with 16-byte bulk copies
    auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(dev);
    const bool use_norm = prm.gamma != nullptr;
    // one allocation for the three gradients: [dW (co x ci) | dgamma (co) | dbeta (co)], handed out as views
    const int64_t co = layout.c_out, ci = layout.c_in;
    Tensor flat = torch::empty({co * ci + 2 * co}, f32);
    Tensor d_w = flat.narrow(0, 0, co * ci).view({co, ci});
    Tensor d_g = use_norm ? flat.narrow(0, co * ci, co) : Tensor();
    Tensor d_b = flat.narrow(0, co * ci + co, co);
    char *base = static_cast<char *>(buf.data_ptr());
    check(rdp_pfn_bwd(points.data_ptr<float>(), points.size(0), &geom, &layout, &prm, base, (size_t)ws_bytes,
                      reinterpret_cast<int32_t *>(base + off_counters), g.data_ptr<float>(), nullptr, argpos.data_ptr<int32_t>(),
                      train_bn ? reinterpret_cast<double *>(base + off_bn) : nullptr, d_w.data_ptr<float>(),
                      use_norm ? d_g.data_ptr<float>() : nullptr, d_b.data_ptr<float>(), at::cuda::getCurrentCUDAStream(dev.index()).stream()),
          "rdp_pfn_bwd");
    return {d_w, d_g, d_b};
}

struct Side {   // one side stream per device for the second encoder of a pair
    static c10::cuda::CUDAStream get(int dev) {
        static std::vector<c10::optional<c10::cuda::CUDAStream>> streams(64);
        static std::mutex m;
        std::lock_guard<std::mutex> lock(m);
        if (!streams[dev].has_value()) streams[dev] = c10::cuda::getStreamFromPool(false, dev);
        return *streams[dev];
    }
};

void stream_wait(const c10::cuda::CUDAStream &waiter, const c10::cuda::CUDAStream &on) {
    static thread_local std::vector<cudaEvent_t> ev(64, nullptr);
    const int dev = on.device_index();
    if (!ev[dev]) C10_CUDA_CHECK(cudaEventCreateWithFlags(&ev[dev], cudaEventDisableTiming));
    C10_CUDA_CHECK(cudaEventRecord(ev[dev], on.stream()));
    C10_CUDA_CHECK(cudaStreamWaitEvent(waiter.stream(), ev[dev], 0));
}

// state of one encoder kept by the autograd node
struct Saved {
    rdp_geom_t geom;
    rdp_layout_t layout;
    double eps, momentum;
    int64_t ws_bytes, off_counters, off_bn;
    bool train_bn, active;
    Tensor points, buf, argpos, weight, bias, gamma, beta, rmean, rvar;
    int slot_w = -1, slot_b = -1, slot_g = -1, slot_be = -1;   // positions of this encoder's parameters among the node's inputs
};
Saved save_of(const Enc &e, const Launched &L, const Finished &f) {
    Saved s;
    s.geom = e.geom; s.layout = e.layout; s.eps = e.eps; s.momentum = e.momentum;
    s.ws_bytes = e.ws_bytes; s.off_counters = e.off_counters; s.off_bn = e.off_bn; s.train_bn = L.train_bn; s.active = e.want_grad;
    s.points = L.points; s.buf = L.buf; s.argpos = f.argpos;
    s.weight = e.weight; s.bias = e.bias; s.gamma = e.gamma; s.beta = e.beta; s.rmean = e.rmean; s.rvar = e.rvar;
    return s;
}
struct PairState : torch::CustomClassHolder {
    Saved a, b;
    Tensor fa, ca, fb, cb;   // the outputs, handed to the autograd node's forward and dropped there (no reference cycle)
};

rdp_pfn_params_t params_of(const Saved &s) {
    rdp_pfn_params_t p;
    memset(&p, 0, sizeof(p));
    auto f = [](const Tensor &t) -> float * { return t.defined() ? t.data_ptr<float>() : nullptr; };
    p.weight = f(s.weight); p.bias = f(s.bias); p.gamma = f(s.gamma); p.beta = f(s.beta);
    p.running_mean = f(s.rmean); p.running_var = f(s.rvar);
    p.eps = s.eps; p.momentum = s.momentum; p.train_bn = s.train_bn ? 1 : 0;
    return p;
}

// The autograd node of the pair.  Inputs: the (up to) four parameters of each encoder; outputs: features / coords of both.
class PairFn : public torch::autograd::Function<PairFn> {
public:
    static torch::autograd::variable_list forward(torch::autograd::AutogradContext *ctx, at::TensorList params,
                                                  c10::intrusive_ptr<PairState> st) {
        ctx->saved_data["state"] = c10::IValue(st);
        ctx->saved_data["n"] = (int64_t)params.size();
        // the features were produced by the kernels already (pair_forward); they leave the state here
        Tensor fa = std::move(st->fa), ca = std::move(st->ca), fb = std::move(st->fb), cb = std::move(st->cb);
        st->fa = Tensor(); st->ca = Tensor(); st->fb = Tensor(); st->cb = Tensor();
        std::vector<Tensor> nd = {ca, cb};
        if (!st->a.active) nd.push_back(fa);
        if (!st->b.active) nd.push_back(fb);
        ctx->mark_non_differentiable(nd);
        ctx->set_materialize_grads(false);
        return {fa, ca, fb, cb};
    }
    static torch::autograd::variable_list backward(torch::autograd::AutogradContext *ctx, torch::autograd::variable_list grads) {
        auto st = ctx->saved_data["state"].toCustomClass<PairState>();
        const Tensor &gfa = grads[0], &gfb = grads[2];
        torch::autograd::variable_list out((size_t)ctx->saved_data["n"].toInt() + 1);   // the parameters, then the state argument
        const int dev = st->a.points.device().index();
        c10::cuda::CUDAGuard guard(dev);
        auto main = at::cuda::getCurrentCUDAStream(dev);
        auto side = Side::get(dev);
        auto place = [&](const Saved &s, const std::vector<Tensor> &g) {   // g = [dW, dgamma | undefined, dbeta_or_dbias]
            const bool norm = s.gamma.defined();
            if (s.slot_w >= 0) out[s.slot_w] = g[0];
            if (norm) {
                if (s.slot_g >= 0) out[s.slot_g] = g[1];
                if (s.slot_be >= 0) out[s.slot_be] = g[2];
            } else if (s.slot_b >= 0) {
                out[s.slot_b] = g[2];
            }
        };
        std::vector<Tensor> gb;
        if (gfb.defined() && st->b.active) {   // the short one goes out first, on the side stream
            stream_wait(side, main);
            c10::cuda::CUDACachingAllocator::recordStream(gfb.storage().data_ptr(), side);
            c10::cuda::CUDAStreamGuard sg(side);
            rdp_pfn_params_t p = params_of(st->b);
            gb = backward_one(st->b.geom, st->b.layout, p, st->b.points, st->b.buf, st->b.ws_bytes, st->b.off_counters, st->b.off_bn,
                              st->b.train_bn, st->b.argpos, gfb);
        }
        if (gfa.defined() && st->a.active) {
            rdp_pfn_params_t p = params_of(st->a);
            place(st->a, backward_one(st->a.geom, st->a.layout, p, st->a.points, st->a.buf, st->a.ws_bytes, st->a.off_counters, st->a.off_bn,
                                      st->a.train_bn, st->a.argpos, gfa));
        }
        if (!gb.empty()) {
            stream_wait(main, side);
            c10::cuda::CUDACachingAllocator::recordStream(gb[0].storage().data_ptr(), main);
            place(st->b, gb);
        }
        return out;
    }
};

// (features_a, coords_a, argpos_a | None, buf_a, n_kept_a, n_pillars_a, the same for b)
py::tuple pair_forward(const py::tuple &ta, const py::tuple &tb) {
    Enc a = unpack(ta), b = unpack(tb);
    const int dev = a.points.device().index();
    TORCH_CHECK(b.points.device().index() == dev, "both encoders must run on the same device");
    Launched La, Lb;
    Finished fa, fb;
    {
        py::gil_scoped_release nogil;
        c10::cuda::CUDAGuard guard(dev);
        auto main = at::cuda::getCurrentCUDAStream(dev);
        auto side = Side::get(dev);
        stream_wait(side, main);   // inputs of `b` were produced on the current stream
        La = launch(a);            // the long kernels go first: they keep the GPU busy while `b` is enqueued
        {
            c10::cuda::CUDAStreamGuard sg(side);
            Lb = launch(b);
        }
        fa = finish(a, La);
        fb = finish(b, Lb);
        stream_wait(main, side);
        c10::cuda::CUDACachingAllocator::recordStream(Lb.features.storage().data_ptr(), main);
        c10::cuda::CUDACachingAllocator::recordStream(Lb.coords.storage().data_ptr(), main);
    }
    Tensor out_fa = fa.features, out_ca = fa.coords, out_fb = fb.features, out_cb = fb.coords;
    if (a.want_grad || b.want_grad) {
        auto st = c10::make_intrusive<PairState>();
        st->a = save_of(a, La, fa);
        st->b = save_of(b, Lb, fb);
        st->fa = fa.features; st->ca = fa.coords; st->fb = fb.features; st->cb = fb.coords;
        torch::autograd::variable_list params;
        auto add = [&](const Tensor &t, int *slot) {
            if (t.defined()) { *slot = (int)params.size(); params.push_back(t); }
        };
        if (a.want_grad) { add(a.weight, &st->a.slot_w); add(a.bias, &st->a.slot_b); add(a.gamma, &st->a.slot_g); add(a.beta, &st->a.slot_be); }
        if (b.want_grad) { add(b.weight, &st->b.slot_w); add(b.bias, &st->b.slot_b); add(b.gamma, &st->b.slot_g); add(b.beta, &st->b.slot_be); }
        auto r = PairFn::apply(at::TensorList(params), st);   // a TensorList argument is what custom functions scan for variables
        out_fa = r[0]; out_ca = r[1]; out_fb = r[2]; out_cb = r[3];
    }
    auto none_or = [](const Tensor &t) -> py::object { return t.defined() ? py::cast(t) : py::none(); };
    return py::make_tuple(out_fa, out_ca, none_or(fa.argpos), La.buf, fa.n_kept, fa.n_pillars, out_fb, out_cb, none_or(fb.argpos), Lb.buf,
                          fb.n_kept, fb.n_pillars);
}

}  // namespace

TORCH_LIBRARY(rdp_host, m) { m.class_<PairState>("PairState"); }

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "native host path of the paired pillar-encoding step (radardistill_b200)";
    m.def("pair_forward", &pair_forward, "enqueue both encoders, wait for (N, P), wire one autograd node");
    m.def("abi_version", []() { return rdp_abi_version(); });
}
