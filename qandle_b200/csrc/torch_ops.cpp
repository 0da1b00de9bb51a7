// Thin torch.library layer over the C ABI (include/qandle_b200.h).  PyTorch is plumbing here: device
// memory from the caching allocator, the current CUDA stream, and autograd registration on the Python
// side (qandle_b200/engine.py).  No compute happens in this file.
//
// Replaces, for the new engine's UnsplittedCircuit.forward, the reference's per-gate python loop
// (reference src/qandle/qcircuit.py:163-174) and the autograd tape over it.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/types.h>

#include <tuple>
#include <vector>

#include "../../include/qandle_b200.h"

namespace {

qb_plan* as_plan(int64_t h) {
  TORCH_CHECK(h != 0, "qandle_b200: null plan handle");
  return reinterpret_cast<qb_plan*>(static_cast<intptr_t>(h));
}

#define QB_CHECK(call)                                                   \
  do {                                                                   \
    int _rc = (call);                                                    \
    TORCH_CHECK(_rc == 0, "qandle_b200: ", #call, " -> ", qb_last_error()); \
  } while (0)

int64_t plan_create(const at::Tensor& program, int64_t n_qubits, int64_t dtype, c10::IntArrayRef opts) {
  TORCH_CHECK(program.device().is_cpu() && program.scalar_type() == at::kInt, "program must be a CPU int32 tensor");
  TORCH_CHECK(program.dim() == 2 && program.size(1) == 4, "program must be [G, 4]");
  at::Tensor prog = program.contiguous();
  qb_plan_opts o{};
  int32_t* of = reinterpret_cast<int32_t*>(&o);
  for (size_t i = 0; i < opts.size() && i < 16; ++i) of[i] = static_cast<int32_t>(opts[i]);
  qb_plan* plan = nullptr;
  QB_CHECK(qb_plan_create(prog.data_ptr<int32_t>(), static_cast<int32_t>(prog.size(0)), static_cast<int32_t>(n_qubits),
                          static_cast<int32_t>(dtype), &o, &plan));
  return static_cast<int64_t>(reinterpret_cast<intptr_t>(plan));
}

void plan_destroy(int64_t h) {
  if (h != 0) qb_plan_destroy(as_plan(h));
}

at::Tensor plan_dump(int64_t h) {
  int64_t n = qb_plan_dump(as_plan(h), nullptr, 0);
  at::Tensor out = at::empty({n}, at::TensorOptions().dtype(at::kLong));
  qb_plan_dump(as_plan(h), out.data_ptr<int64_t>(), n);
  return out;
}

at::Tensor plan_info(int64_t h) {
  // [num_steps, num_sweeps, num_groups, launches_fwd_probs, launches_bwd]
  qb_plan* p = as_plan(h);
  at::Tensor out = at::empty({5}, at::TensorOptions().dtype(at::kLong));
  int64_t* o = out.data_ptr<int64_t>();
  o[0] = qb_plan_num_steps(p);
  o[1] = qb_plan_num_sweeps(p);
  o[2] = qb_plan_num_groups(p);
  o[3] = qb_plan_num_launches(p, 0, QB_MEASURE_PROBS);
  o[4] = qb_plan_num_launches(p, 1, QB_MEASURE_PROBS);
  return out;
}

int64_t workspace_bytes(int64_t h, int64_t batch) { return qb_workspace_bytes(as_plan(h), batch); }

const void* ptr_or_null(const at::Tensor& t) { return t.defined() && t.numel() > 0 ? t.data_ptr() : nullptr; }

void check_cuda_contig(const at::Tensor& t, const char* name) {
  if (!t.defined() || t.numel() == 0) return;
  TORCH_CHECK(t.is_cuda(), "qandle_b200: ", name, " must be a CUDA tensor (there is no CPU fallback)");
  TORCH_CHECK(t.is_contiguous(), "qandle_b200: ", name, " must be contiguous");
}

at::Tensor make_workspace(int64_t h, int64_t batch, const at::Device& dev) {
  int64_t bytes = qb_workspace_bytes(as_plan(h), batch);
  TORCH_CHECK(bytes >= 0, "qandle_b200: workspace size query failed");
  return at::empty({bytes + 256}, at::TensorOptions().dtype(at::kByte).device(dev));
}

void* aligned(const at::Tensor& ws) {
  uintptr_t p = reinterpret_cast<uintptr_t>(ws.data_ptr());
  return reinterpret_cast<void*>((p + 255) & ~uintptr_t(255));
}

// ---- one-call forward / backward (unsharded) ---------------------------------------------------------
std::tuple<at::Tensor, at::Tensor> circuit_forward(int64_t h, const at::Tensor& shared_angles, const at::Tensor& batch_angles,
                                                   const at::Tensor& fixed_mats, const c10::optional<at::Tensor>& init_state,
                                                   int64_t batch, int64_t n_qubits, int64_t measure) {
  TORCH_CHECK(shared_angles.is_cuda(), "qandle_b200: tensors must live on a CUDA device (there is no CPU fallback)");
  c10::cuda::CUDAGuard guard(shared_angles.device());
  check_cuda_contig(shared_angles, "shared_angles");
  check_cuda_contig(batch_angles, "batch_angles");
  check_cuda_contig(fixed_mats, "fixed_mats");
  const bool dbl = shared_angles.scalar_type() == at::kDouble;
  TORCH_CHECK(dbl || shared_angles.scalar_type() == at::kFloat, "angles must be float32 or float64");
  const auto cdtype = dbl ? at::kComplexDouble : at::kComplexFloat;
  const int64_t N = int64_t(1) << n_qubits;
  at::Tensor state;
  int init_kind = QB_INIT_ZERO;
  if (init_state.has_value() && init_state->defined()) {
    TORCH_CHECK(init_state->scalar_type() == cdtype, "init_state dtype does not match the angle dtype");
    TORCH_CHECK(init_state->dim() == 2 && init_state->size(0) == batch && init_state->size(1) == N, "init_state must be [B, 2^n]");
    state = init_state->clone(at::MemoryFormat::Contiguous);
    init_kind = QB_INIT_STATE;
  } else {
    state = at::empty({batch, N}, shared_angles.options().dtype(cdtype));
  }
  at::Tensor ws = make_workspace(h, batch, shared_angles.device());
  at::Tensor out;
  void* out_ptr = nullptr;
  if (measure == QB_MEASURE_PROBS) {
    out = at::empty({batch, n_qubits}, shared_angles.options());
    out_ptr = out.data_ptr();
  } else if (measure == QB_MEASURE_JOINT) {
    out = at::empty({batch, N}, shared_angles.options());
    out_ptr = out.data_ptr();
  }
  const int32_t ncols = batch_angles.defined() && batch_angles.dim() == 2 ? static_cast<int32_t>(batch_angles.size(1)) : 0;
  auto stream = at::cuda::getCurrentCUDAStream();
  QB_CHECK(qb_forward_dev(as_plan(h), batch, ptr_or_null(shared_angles), ptr_or_null(batch_angles), ncols,
                          ptr_or_null(fixed_mats), init_kind, state.data_ptr(), static_cast<int32_t>(measure), out_ptr,
                          aligned(ws), stream.stream()));
  // (out, final state in the internal layout -- saved for the adjoint backward).  MeasureState: the state IS the result; the
  // second output is then empty (an op must not return the same tensor twice).
  if (measure == QB_MEASURE_STATE) return std::make_tuple(state, at::empty({0}, state.options()));
  return std::make_tuple(out, state);
}

// Adjoint-state backward.  `state` is the forward's second output (MeasureState: its first) and is only READ: the un-computed
// copy lives in a scratch buffer, so autograd may call this any number of times for one forward (retain_graph, double use).
std::tuple<at::Tensor, at::Tensor, at::Tensor> circuit_backward(int64_t h, const at::Tensor& shared_angles,
                                                                const at::Tensor& batch_angles, const at::Tensor& fixed_mats,
                                                                const at::Tensor& state, const at::Tensor& grad_out, int64_t measure,
                                                                bool want_init_grad) {
  TORCH_CHECK(state.is_cuda(), "qandle_b200: tensors must live on a CUDA device (there is no CPU fallback)");
  c10::cuda::CUDAGuard guard(state.device());
  check_cuda_contig(state, "state");
  check_cuda_contig(grad_out, "grad_out");
  check_cuda_contig(shared_angles, "shared_angles");
  check_cuda_contig(batch_angles, "batch_angles");
  const int64_t batch = state.size(0);
  at::Tensor work = at::empty_like(state);
  at::Tensor lam = at::empty_like(state);
  at::Tensor ws = make_workspace(h, batch, state.device());
  at::Tensor g_shared = at::empty_like(shared_angles);  // (qb_finalize_grads_dev overwrites both gradient buffers)
  at::Tensor g_batch = batch_angles.defined() ? at::empty_like(batch_angles) : at::Tensor();
  const int32_t ncols = batch_angles.defined() && batch_angles.dim() == 2 ? static_cast<int32_t>(batch_angles.size(1)) : 0;
  auto stream = at::cuda::getCurrentCUDAStream();
  QB_CHECK(qb_backward_from_dev(as_plan(h), batch, ptr_or_null(shared_angles), ptr_or_null(batch_angles), ncols,
                                ptr_or_null(fixed_mats), state.data_ptr(), work.data_ptr(), lam.data_ptr(), static_cast<int32_t>(measure),
                                grad_out.data_ptr(), g_shared.numel() ? g_shared.data_ptr() : nullptr,
                                static_cast<int32_t>(g_shared.numel()), g_batch.defined() && g_batch.numel() ? g_batch.data_ptr() : nullptr,
                                aligned(ws), stream.stream()));
  at::Tensor g_init;
  if (want_init_grad) {
    QB_CHECK(qb_convert_layout_dev(as_plan(h), batch, lam.data_ptr(), stream.stream()));  // internal -> interleaved
    g_init = lam.mul_(2);  // torch convention: grad = 2 dL/dpsi0*
  }
  else g_init = at::empty({0}, state.options());
  if (!g_batch.defined()) g_batch = at::empty({0}, shared_angles.options());
  return std::make_tuple(g_shared, g_batch, g_init);
}

// ---- step-level ops for the amplitude-sharded driver (qandle_b200/distributed.py) --------------------
void prepare(int64_t h, int64_t batch, const at::Tensor& shared_angles, const at::Tensor& batch_angles,
             const at::Tensor& fixed_mats, at::Tensor workspace) {
  c10::cuda::CUDAGuard guard(workspace.device());
  const int32_t ncols = batch_angles.defined() && batch_angles.dim() == 2 ? static_cast<int32_t>(batch_angles.size(1)) : 0;
  QB_CHECK(qb_prepare_dev(as_plan(h), batch, ptr_or_null(shared_angles), ptr_or_null(batch_angles), ncols,
                          ptr_or_null(fixed_mats), aligned(workspace), at::cuda::getCurrentCUDAStream().stream()));
}

void init_zero(int64_t h, int64_t batch, at::Tensor state, int64_t rank) {
  c10::cuda::CUDAGuard guard(state.device());
  QB_CHECK(qb_init_zero_dev(as_plan(h), batch, state.data_ptr(), static_cast<int32_t>(rank),
                            at::cuda::getCurrentCUDAStream().stream()));
}

void apply_forward(int64_t h, int64_t step_begin, int64_t step_end, int64_t batch, at::Tensor state, at::Tensor workspace,
                   int64_t rank) {
  c10::cuda::CUDAGuard guard(state.device());
  check_cuda_contig(state, "state");
  QB_CHECK(qb_apply_forward_dev(as_plan(h), static_cast<int32_t>(step_begin), static_cast<int32_t>(step_end), batch,
                                state.data_ptr(), aligned(workspace), static_cast<int32_t>(rank),
                                at::cuda::getCurrentCUDAStream().stream()));
}

void apply_backward(int64_t h, int64_t step_begin, int64_t step_end, int64_t batch, at::Tensor state, at::Tensor lam,
                    at::Tensor workspace, int64_t rank) {
  c10::cuda::CUDAGuard guard(state.device());
  check_cuda_contig(state, "state");
  check_cuda_contig(lam, "lambda");
  QB_CHECK(qb_apply_backward_dev(as_plan(h), static_cast<int32_t>(step_begin), static_cast<int32_t>(step_end), batch,
                                 state.data_ptr(), lam.data_ptr(), aligned(workspace), static_cast<int32_t>(rank),
                                 at::cuda::getCurrentCUDAStream().stream()));
}

at::Tensor measure_probs(int64_t h, int64_t batch, int64_t n_qubits, const at::Tensor& state, at::Tensor workspace, int64_t rank) {
  c10::cuda::CUDAGuard guard(state.device());
  check_cuda_contig(state, "state");
  at::Tensor out = at::empty({batch, n_qubits}, state.options().dtype(state.scalar_type() == at::kComplexDouble ? at::kDouble : at::kFloat));
  QB_CHECK(qb_measure_probs_dev(as_plan(h), batch, state.data_ptr(), out.data_ptr(), aligned(workspace),
                                static_cast<int32_t>(rank), at::cuda::getCurrentCUDAStream().stream()));
  return out;
}

void seed_probs(int64_t h, int64_t batch, const at::Tensor& state, const at::Tensor& grad, at::Tensor lam, int64_t rank) {
  c10::cuda::CUDAGuard guard(state.device());
  check_cuda_contig(grad, "grad");
  QB_CHECK(qb_seed_probs_dev(as_plan(h), batch, state.data_ptr(), grad.data_ptr(), lam.data_ptr(), static_cast<int32_t>(rank),
                             at::cuda::getCurrentCUDAStream().stream()));
}

void backward_begin(int64_t h, int64_t batch, at::Tensor workspace) {
  c10::cuda::CUDAGuard guard(workspace.device());
  QB_CHECK(qb_backward_begin_dev(as_plan(h), batch, aligned(workspace), at::cuda::getCurrentCUDAStream().stream()));
}

std::tuple<at::Tensor, at::Tensor> finalize_grads(int64_t h, int64_t batch, const at::Tensor& shared_angles,
                                                  const at::Tensor& batch_angles, const at::Tensor& fixed_mats,
                                                  at::Tensor workspace) {
  c10::cuda::CUDAGuard guard(workspace.device());
  at::Tensor g_shared = at::zeros_like(shared_angles);
  at::Tensor g_batch = batch_angles.defined() ? at::zeros_like(batch_angles) : at::empty({0}, shared_angles.options());
  const int32_t ncols = batch_angles.defined() && batch_angles.dim() == 2 ? static_cast<int32_t>(batch_angles.size(1)) : 0;
  QB_CHECK(qb_finalize_grads_dev(as_plan(h), batch, ptr_or_null(shared_angles), ptr_or_null(batch_angles), ncols,
                                 ptr_or_null(fixed_mats), aligned(workspace), g_shared.numel() ? g_shared.data_ptr() : nullptr,
                                 static_cast<int32_t>(g_shared.numel()), g_batch.numel() ? g_batch.data_ptr() : nullptr,
                                 at::cuda::getCurrentCUDAStream().stream()));
  return std::make_tuple(g_shared, g_batch);
}

void exchange_p2p(int64_t h, int64_t batch, at::Tensor state, c10::IntArrayRef peer_ptrs, int64_t rank, int64_t step) {
  c10::cuda::CUDAGuard guard(state.device());
  std::vector<const void*> pp(peer_ptrs.size());
  for (size_t i = 0; i < peer_ptrs.size(); ++i) pp[i] = reinterpret_cast<const void*>(static_cast<intptr_t>(peer_ptrs[i]));
  TORCH_CHECK(reinterpret_cast<const void*>(state.data_ptr()) == pp[rank], "qandle_b200: state is not the local symmetric buffer");
  QB_CHECK(qb_exchange_p2p_dev(as_plan(h), batch, pp.data(), static_cast<int32_t>(rank), static_cast<int32_t>(pp.size()),
                               static_cast<int32_t>(step), at::cuda::getCurrentCUDAStream().stream()));
}

void exchange_push(int64_t h, int64_t batch, at::Tensor state, c10::IntArrayRef staging_ptrs, int64_t rank, int64_t piece, int64_t pieces,
                   int64_t phase) {
  c10::cuda::CUDAGuard guard(state.device());
  std::vector<const void*> pp(staging_ptrs.size());
  for (size_t i = 0; i < staging_ptrs.size(); ++i) pp[i] = reinterpret_cast<const void*>(static_cast<intptr_t>(staging_ptrs[i]));
  QB_CHECK(qb_exchange_push_dev(as_plan(h), batch, state.data_ptr(), pp.data(), static_cast<int32_t>(rank), static_cast<int32_t>(pp.size()),
                                static_cast<int32_t>(piece), static_cast<int32_t>(pieces), static_cast<int32_t>(phase),
                                at::cuda::getCurrentCUDAStream().stream()));
}

}  // namespace

TORCH_LIBRARY(qandle_b200, m) {
  m.def("plan_create(Tensor program, int n_qubits, int dtype, int[] opts) -> int", &plan_create);
  m.def("plan_destroy(int plan) -> ()", &plan_destroy);
  m.def("plan_dump(int plan) -> Tensor", &plan_dump);
  m.def("plan_info(int plan) -> Tensor", &plan_info);
  m.def("workspace_bytes(int plan, int batch) -> int", &workspace_bytes);
  m.def(
      "circuit_forward(int plan, Tensor shared_angles, Tensor batch_angles, Tensor fixed_mats, Tensor? init_state, int "
      "batch, int n_qubits, int measure) -> (Tensor, Tensor)");
  m.def(
      "circuit_backward(int plan, Tensor shared_angles, Tensor batch_angles, Tensor fixed_mats, Tensor state, Tensor "
      "grad_out, int measure, bool want_init_grad) -> (Tensor, Tensor, Tensor)");
  m.def("prepare(int plan, int batch, Tensor shared_angles, Tensor batch_angles, Tensor fixed_mats, Tensor(a!) workspace) -> ()");
  m.def("init_zero(int plan, int batch, Tensor(a!) state, int rank) -> ()");
  m.def("apply_forward(int plan, int step_begin, int step_end, int batch, Tensor(a!) state, Tensor(b!) workspace, int rank) -> ()");
  m.def(
      "apply_backward(int plan, int step_begin, int step_end, int batch, Tensor(a!) state, Tensor(b!) lam, Tensor(c!) "
      "workspace, int rank) -> ()");
  m.def("measure_probs(int plan, int batch, int n_qubits, Tensor state, Tensor(a!) workspace, int rank) -> Tensor");
  m.def("exchange_p2p(int plan, int batch, Tensor(a!) state, int[] peer_ptrs, int rank, int step) -> ()");
  m.def("exchange_push(int plan, int batch, Tensor(a!) state, int[] staging_ptrs, int rank, int piece, int pieces, int phase) -> ()");
  m.def("seed_probs(int plan, int batch, Tensor state, Tensor grad, Tensor(a!) lam, int rank) -> ()");
  m.def("backward_begin(int plan, int batch, Tensor(a!) workspace) -> ()");
  m.def(
      "finalize_grads(int plan, int batch, Tensor shared_angles, Tensor batch_angles, Tensor fixed_mats, Tensor(a!) "
      "workspace) -> (Tensor, Tensor)");
}

TORCH_LIBRARY_IMPL(qandle_b200, CUDA, m) {
  m.impl("circuit_forward", &circuit_forward);
  m.impl("circuit_backward", &circuit_backward);
  m.impl("prepare", &prepare);
  m.impl("init_zero", &init_zero);
  m.impl("apply_forward", &apply_forward);
  m.impl("apply_backward", &apply_backward);
  m.impl("measure_probs", &measure_probs);
  m.impl("exchange_p2p", &exchange_p2p);
  m.impl("exchange_push", &exchange_push);
  m.impl("seed_probs", &seed_probs);
  m.impl("backward_begin", &backward_begin);
  m.impl("finalize_grads", &finalize_grads);
}
