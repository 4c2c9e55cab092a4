// C-ABI library of the B200-native FDM + Parareal hot path.
//
// * plans: NVRTC-compiles the generated stage kernels (fdm_template.cuh behind a
//   generated prelude) for sm_100a, loads the cubin through the driver API
//   (entry points fetched with cudaGetDriverEntryPoint, so the library does not
//   link libcuda and can be dlopen'ed on a machine without a driver) and runs
//   the device-resident time loop of FDMOperator.solve
//   (reference fdm_operator.py:48-165, numerical_integrator.py:47-132);
// * fixed sm_100a kernels: layout conversion and the Parareal state updates
//   (reference parareal_operator.py:164, 183-185, 85-100, 192).
//
// See include/pararealml_b200.h for the contract of every entry point.
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvrtc.h>

#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../../include/pararealml_b200.h"

namespace {

thread_local std::string g_error;

int fail(const std::string& msg) {
  g_error = msg;
  return -1;
}

#define PML_CUDA(call)                                                       \
  do {                                                                       \
    cudaError_t e_ = (call);                                                 \
    if (e_ != cudaSuccess)                                                   \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_));       \
  } while (0)

// ---- driver API through the runtime (no libcuda link dependency) ----------
struct Driver {
  CUresult (*moduleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*moduleUnload)(CUmodule) = nullptr;
  CUresult (*moduleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*launchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned,
                           unsigned, unsigned, unsigned, CUstream, void**,
                           void**) = nullptr;
  CUresult (*launchCooperativeKernel)(CUfunction, unsigned, unsigned, unsigned,
                                      unsigned, unsigned, unsigned, unsigned,
                                      CUstream, void**) = nullptr;
  CUresult (*occupancyMaxActiveBlocks)(int*, CUfunction, int, size_t) = nullptr;
  CUresult (*getErrorString)(CUresult, const char**) = nullptr;
  CUresult (*funcSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*tensorMapEncodeTiled)(
      CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
      CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
      CUtensorMapFloatOOBfill) = nullptr;
  bool ready = false;
};
Driver g_drv;

int load_driver() {
  if (g_drv.ready) return 0;
  PML_CUDA(cudaFree(0));  // makes the primary context current
  struct {
    const char* name;
    void** slot;
  } syms[] = {
      {"cuModuleLoadData", (void**)&g_drv.moduleLoadData},
      {"cuModuleUnload", (void**)&g_drv.moduleUnload},
      {"cuModuleGetFunction", (void**)&g_drv.moduleGetFunction},
      {"cuLaunchKernel", (void**)&g_drv.launchKernel},
      {"cuLaunchCooperativeKernel", (void**)&g_drv.launchCooperativeKernel},
      {"cuOccupancyMaxActiveBlocksPerMultiprocessor",
       (void**)&g_drv.occupancyMaxActiveBlocks},
      {"cuGetErrorString", (void**)&g_drv.getErrorString},
      {"cuFuncSetAttribute", (void**)&g_drv.funcSetAttribute},
      {"cuTensorMapEncodeTiled", (void**)&g_drv.tensorMapEncodeTiled},
  };
  for (auto& s : syms) {
    cudaDriverEntryPointQueryResult q;
    PML_CUDA(cudaGetDriverEntryPoint(s.name, s.slot, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || *s.slot == nullptr)
      return fail(std::string("driver entry point not found: ") + s.name);
  }
  g_drv.ready = true;
  return 0;
}

std::string cu_err(CUresult r) {
  const char* s = nullptr;
  if (g_drv.getErrorString) g_drv.getErrorString(r, &s);
  return s ? s : "unknown driver error";
}

#define PML_CU(call)                                                         \
  do {                                                                       \
    CUresult r_ = (call);                                                    \
    if (r_ != CUDA_SUCCESS) return fail(std::string(#call) + ": " + cu_err(r_)); \
  } while (0)

// internal: two forward Euler steps in one stage-pair launch
#define PML_INTEGRATOR_FE_PAIR 100

// cells along axis 0 per thread of the Jacobi sweep (codegen.py JACOBI_REP)
#define PML_JACOBI_REP 4

// mirror of PmlArgs in fdm_template.cuh (kept layout-identical)
struct PmlArgs {
  const double* u;
  const double* y;
  const double* acc_in;
  double* u_out;
  double* acc_out;
  double* y_next;
  double* lap_rhs;
  double t_eval;
  double dt;
  const double* neu[6];
  const double* dir[6];
  const double* dir_full[6];
  const double* coord[3];
  const double* aux[4];
};

struct PmlFusedArgs {
  alignas(64) CUtensorMap tm_in;
  alignas(64) CUtensorMap tm_y;
  alignas(64) CUtensorMap tm_acc;
  PmlArgs s;
  double t_eval_b;
  const double* neu_b[6];
  const double* dir_b[6];
  int z_begin, z_end;
};

struct PmlSmallArgs {
  PmlArgs s;
  long long neu_stride[6];
  long long dir_stride[6];
  double* traj;
  long long stride;
  const double* t;
  int n_steps;
  int integrator;
  long long slot0;
  double* u_a;
  double* u_b;
  double* acc;
  long long y_batch_stride;
  long long traj_batch_stride;
  long long ws_batch_stride;
};

struct PmlJacobiArgs {
  PmlArgs base;
  const double* y_hat;
  const double* rhs;
  double* y_new;
  double* partials;
  int* flags;
  double tol;
};

// `name` is the file name recorded in the line info; when it is a real path the
// generated source is also written there so that profilers can show it
int nvrtc_compile(const char* source, std::vector<char>& cubin,
                  const char* name = "pml_generated.cu") {
  nvrtcProgram prog;
  if (nvrtcCreateProgram(&prog, source, name, 0, nullptr, nullptr) !=
      NVRTC_SUCCESS)
    return fail("nvrtcCreateProgram failed");
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17",
                        "-lineinfo", "--fmad=true"};
  nvrtcResult r = nvrtcCompileProgram(prog, 4, opts);
  if (r != NVRTC_SUCCESS) {
    size_t n = 0;
    nvrtcGetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    nvrtcGetProgramLog(prog, &log[0]);
    nvrtcDestroyProgram(&prog);
    return fail(std::string("NVRTC: ") + nvrtcGetErrorString(r) + "\n" + log);
  }
  size_t n = 0;
  if (nvrtcGetCUBINSize(prog, &n) != NVRTC_SUCCESS || n == 0) {
    nvrtcDestroyProgram(&prog);
    return fail("NVRTC produced no cubin");
  }
  cubin.resize(n);
  nvrtcGetCUBIN(prog, cubin.data());
  nvrtcDestroyProgram(&prog);
  return 0;
}

bool read_file(const char* path, std::vector<char>& out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  out.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
  return !out.empty();
}

// atomic publish: every writer stages through its own temporary name (ranks
// that compile the same kernel at the same time must not share one inode)
bool write_file(const char* path, const std::vector<char>& data) {
  static unsigned counter = 0;
  std::string tmp = std::string(path) + ".tmp" + std::to_string((long long)getpid()) +
                    "." + std::to_string(counter++);
  {
    std::ofstream f(tmp, std::ios::binary);
    if (!f) return false;
    f.write(data.data(), (std::streamsize)data.size());
  }
  return std::rename(tmp.c_str(), path) == 0;
}

}  // namespace

struct pml_plan {
  pml_plan_desc desc;
  CUmodule module = nullptr;
  CUfunction stage[7] = {};
  // rk4 1+2, rk4 3+4, midpoint 1+2, two forward Euler steps (optional)
  CUfunction fused[4] = {};
  dim3 fgrid, fblock;
  unsigned fsmem[4] = {0, 0, 0, 0};
  CUfunction small_run = nullptr;
  CUfunction eval_rhs = nullptr;
  CUfunction apply_dir = nullptr;
  CUfunction jac_init = nullptr, jac_sweep = nullptr, jac_store = nullptr;
  CUfunction jac_loop = nullptr;  // persistent cooperative form of the sweeps
  int jac_loop_blocks = 0;        // thread blocks resident at once
  pml_tables tables{};
  dim3 grid, block;
  dim3 sgrid;  // grid of the stage kernels (zrep cells along axis 0 per thread)
  dim3 jgrid;  // grid of the Jacobi sweep (PML_JACOBI_REP cells along axis 0)
  long long n_cells = 0;
  long long n_blocks = 0;
  long long launches = 0;
};

namespace {

// boundary tables at their time slots: the kernels receive ready pointers
void bind_neu(const pml_plan* p, const double** neu, long long slot) {
  for (int f = 0; f < 6; ++f)
    neu[f] = p->tables.neu[f]
                 ? p->tables.neu[f] + slot * p->tables.neu_stride[f] : nullptr;
}
void bind_dir(const pml_plan* p, const double** dir, long long slot) {
  for (int f = 0; f < 6; ++f)
    dir[f] = p->tables.dir[f]
                 ? p->tables.dir[f] + slot * p->tables.dir_stride[f] : nullptr;
}
// TMA descriptor of a state (component planes) seen as the 4-D array
// [component][axis 0][axis 1][contiguous axis] with a box of one plane tile
int state_tensor_map(const pml_plan* p, const double* base, unsigned box_x,
                     unsigned box_y, CUtensorMap* out) {
  const pml_plan_desc& d = p->desc;
  const bool three = d.n_dims == 3;
  const cuuint64_t nx = (cuuint64_t)(three ? d.shape[2] : d.shape[1]);
  const cuuint64_t ny = (cuuint64_t)(three ? d.shape[1] : 1);
  const cuuint64_t nz = (cuuint64_t)d.shape[0];
  const cuuint64_t dims[4] = {nx, ny, nz, (cuuint64_t)d.y_dim};
  const cuuint64_t strides[3] = {nx * 8, nx * ny * 8, nx * ny * nz * 8};
  const cuuint32_t box[4] = {box_x, box_y, 1, 1};
  const cuuint32_t elem[4] = {1, 1, 1, 1};
  CUresult r = g_drv.tensorMapEncodeTiled(
      out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void*)base, dims, strides, box,
      elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled: " + cu_err(r));
  return 0;
}

void fill_tables(const pml_plan* p, PmlArgs& a) {
  bind_neu(p, a.neu, 0);
  bind_dir(p, a.dir, 0);
  bind_dir(p, a.dir_full, 0);
  for (int i = 0; i < 3; ++i) a.coord[i] = p->tables.coord[i];
  for (int i = 0; i < 4; ++i) a.aux[i] = p->tables.aux[i];
}

int launch(pml_plan* p, CUfunction fn, void** params, CUstream s) {
  PML_CU(g_drv.launchKernel(fn, p->grid.x, p->grid.y, p->grid.z, p->block.x,
                            p->block.y, p->block.z, 0, s, params, nullptr));
  p->launches += 1;
  return 0;
}

int jacobi_solve(pml_plan* p, const pml_workspace* ws, const PmlArgs& tbl,
                 const double* rhs, const double* y_init, double* y_next,
                 double tol, long long max_sweeps, int* sweeps_out,
                 CUstream s) {
  // tbl carries the slots of t + dt for both Neumann and Dirichlet tables
  PML_CUDA(cudaMemsetAsync(ws->flags, 0, 4 * sizeof(int), (cudaStream_t)s));
  {
    PmlArgs a = tbl;
    const double* init = y_init;
    double* out = ws->jac_a;
    void* params[] = {&a, &init, &out};
    if (launch(p, p->jac_init, params, s)) return -1;
  }
  double* bufs[2] = {ws->jac_a, ws->jac_b};
  int host_flags[2] = {0, 0};
  // all sweeps and their convergence tests in one cooperative launch of a
  // persistent grid (PML_JACOBI_LOOP=0: one launch per sweep, in batches
  // between which the host reads the convergence flag)
  static const bool use_loop = [] {
    const char* e = std::getenv("PML_JACOBI_LOOP");
    return !(e && e[0] == '0');
  }();
  if (use_loop) {
    if (p->jac_loop_blocks == 0) {
      int per_sm = 0, sms = 0, dev = 0;
      PML_CUDA(cudaGetDevice(&dev));
      PML_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      PML_CU(g_drv.occupancyMaxActiveBlocks(
          &per_sm, p->jac_loop, (int)(p->block.x * p->block.y * p->block.z), 0));
      if (per_sm < 1) return fail("the Jacobi loop kernel does not fit an SM");
      p->jac_loop_blocks = per_sm * sms;
    }
    const long long tiles = (long long)p->jgrid.x * p->jgrid.y * p->jgrid.z;
    // two partial sums per block (sweep parity) live in ws->partials
    long long blocks = std::min<long long>(tiles, p->jac_loop_blocks);
    blocks = std::min<long long>(blocks, std::max<long long>(p->n_blocks / 2, 1));
    PmlJacobiArgs j;
    j.base = tbl;
    j.y_hat = bufs[0];
    j.rhs = rhs;
    j.y_new = bufs[1];
    j.partials = ws->partials;
    j.flags = ws->flags;
    j.tol = tol;
    int cap = max_sweeps > 0x7fffffffLL ? 0x7fffffff : (int)max_sweeps;
    int gx = (int)p->jgrid.x, gy = (int)p->jgrid.y, gz = (int)p->jgrid.z;
    void* params[] = {&j, &cap, &gx, &gy, &gz};
    PML_CU(g_drv.launchCooperativeKernel(p->jac_loop, (unsigned)blocks, 1, 1,
                                         p->block.x, p->block.y, p->block.z, 0, s,
                                         params));
    p->launches += 1;
    PML_CUDA(cudaMemcpyAsync(host_flags, ws->flags, 2 * sizeof(int),
                             cudaMemcpyDeviceToHost, (cudaStream_t)s));
    PML_CUDA(cudaStreamSynchronize((cudaStream_t)s));
  } else {
  long long issued = 0;
  // sweeps are enqueued in growing batches; after each batch the host reads
  // the convergence flag (launches after convergence return at once)
  long long batch = 32;
  while (true) {
    long long todo = batch;
    if (max_sweeps > 0 && issued + todo > max_sweeps) todo = max_sweeps - issued;
    if (todo <= 0) break;
    for (long long k = 0; k < todo; ++k, ++issued) {
      PmlJacobiArgs j;
      j.base = tbl;
      j.y_hat = bufs[issued & 1];
      j.rhs = rhs;
      j.y_new = bufs[(issued + 1) & 1];
      j.partials = ws->partials;
      j.flags = ws->flags;
      j.tol = tol;
      void* params[] = {&j};
      PML_CU(g_drv.launchKernel(p->jac_sweep, p->jgrid.x, p->jgrid.y, p->jgrid.z,
                                p->block.x, p->block.y, p->block.z, 0, s, params,
                                nullptr));
      p->launches += 1;
    }
    PML_CUDA(cudaMemcpyAsync(host_flags, ws->flags, 2 * sizeof(int),
                             cudaMemcpyDeviceToHost, (cudaStream_t)s));
    PML_CUDA(cudaStreamSynchronize((cudaStream_t)s));
    if (host_flags[0]) break;
    if (batch < 8192) batch *= 2;
  }
  }
  const long long sweeps = host_flags[1];
  if (sweeps_out) *sweeps_out = (int)sweeps;
  const double* result = bufs[sweeps & 1];
  void* params[] = {&result, &y_next};
  return launch(p, p->jac_store, params, s);
}

// One explicit time step y -> y_next as a sequence of launches ("phases").
// phase < 0 runs all of them; otherwise only that one, and *fresh receives the
// buffer it has written: the next phase's stencil input, or y_next after the
// last phase (a domain-decomposed caller exchanges its halo planes then).
int run_step(pml_plan* p, int integrator, const pml_workspace* ws, PmlArgs& a,
             const double* y, double* y_next, double t, double d_t,
             long long s_t, int phase, double** fresh, CUstream s,
             int z_begin = 0, int z_end = -1, double* y_mid = nullptr,
             double t_next = 0.0) {
  // planes [z_begin, z_end) of axis 0 (stage-pair kernels only); default: all
  const int n_planes = p->desc.shape[0];
  if (z_end < 0) z_end = n_planes;
  const bool whole = z_begin == 0 && z_end == n_planes;
  if (z_begin < 0 || z_end > n_planes || z_begin >= z_end)
    return fail("plane range out of bounds");
  const double half = d_t / 2.0;
  const long long s_h = s_t + 1, s_f = s_t + 2;
  a.y = y;
  a.y_next = y_next;
  bind_dir(p, a.dir_full, s_f);
  void* params[] = {&a};
  int index = 0;  // phase counter
  auto wanted = [&](double* written) {
    const bool run = phase < 0 || phase == index;
    if (run && fresh) *fresh = written;
    ++index;
    return run;
  };
  auto stage = [&](int k, const double* u, double* u_out, double t_eval,
                   long long neu_slot, long long dir_slot) {
    if (!wanted(u_out ? u_out : y_next)) return 0;
    a.u = u;
    a.u_out = u_out;
    a.acc_in = ws->acc;
    a.acc_out = ws->acc;
    a.t_eval = t_eval;
    bind_neu(p, a.neu, neu_slot);
    bind_dir(p, a.dir, dir_slot);
    CUresult r_ = g_drv.launchKernel(p->stage[k], p->sgrid.x, p->sgrid.y,
                                     p->sgrid.z, p->block.x, p->block.y,
                                     p->block.z, 0, s, params, nullptr);
    if (r_ != CUDA_SUCCESS) return fail("stage launch: " + cu_err(r_));
    p->launches += 1;
    return 0;
  };
  auto fused = [&](int k, const double* u, double* u_out, double t_a,
                   long long neu_a, long long dir_a, double t_b,
                   long long neu_b, long long dir_b) {
    if (!wanted(u_out ? u_out : y_next)) return 0;
    PmlFusedArgs f;
    std::memset(&f, 0, sizeof(f));
    const unsigned ftx = (unsigned)p->desc.fused_tile[0];
    const unsigned fty = (unsigned)p->desc.fused_tile[1];
    const unsigned hy = p->desc.n_dims == 3 ? 1u : 0u;
    // stage A's stencil input (tile + halo 2); for stages 3+4 also the
    // step-start state (rows of the stage-A tile) and the accumulator
    if (state_tensor_map(p, u, ftx + 4, fty + 4 * hy, &f.tm_in)) return -1;
    if (k == 1) {
      if (state_tensor_map(p, y, ftx + 4, fty + 2 * hy, &f.tm_y)) return -1;
      if (state_tensor_map(p, ws->acc, ftx, fty, &f.tm_acc)) return -1;
    }
    a.u = u;
    a.u_out = u_out;
    a.acc_in = ws->acc;
    a.acc_out = ws->acc;
    a.t_eval = t_a;
    bind_neu(p, a.neu, neu_a);
    bind_dir(p, a.dir, dir_a);
    f.s = a;
    f.t_eval_b = t_b;
    bind_neu(p, f.neu_b, neu_b);
    bind_dir(p, f.dir_b, dir_b);
    f.z_begin = z_begin;
    f.z_end = z_end;
    // the chunks of the marching axis are the last used grid dimension
    const unsigned chunks =
        (unsigned)((z_end - z_begin + p->desc.fused_zc - 1) / p->desc.fused_zc);
    dim3 grid = p->fgrid;
    if (p->desc.n_dims == 3) grid.z = chunks; else grid.y = chunks;
    void* fparams[] = {&f};
    CUresult r_ = g_drv.launchKernel(p->fused[k], grid.x, grid.y,
                                     grid.z, p->fblock.x, p->fblock.y, 1,
                                     p->fsmem[k], s, fparams, nullptr);
    if (r_ != CUDA_SUCCESS) return fail("fused launch: " + cu_err(r_));
    p->launches += 1;
    return 0;
  };
  int rc = 0;
  // the fused kernels move boxes with the TMA unit: 16-byte aligned planes.
  // (phase-wise callers get the fused sequence whenever the plan has it: the
  // number of phases must not depend on pointer values)
  auto aligned = [](const void* q) { return ((uintptr_t)q & 15u) == 0; };
  const bool all_aligned = aligned(y) && aligned(y_next) && aligned(ws->u_b) &&
                           aligned(ws->acc);
  if (p->desc.fused && phase >= 0 && !all_aligned)
    return fail("phase-wise stepping needs 16-byte aligned states");
  const bool use_fused = p->desc.fused && all_aligned;
  if (!whole && (!use_fused || integrator == PML_INTEGRATOR_FORWARD_EULER))
    return fail("plane ranges need the stage-pair kernels");
  if (integrator == PML_INTEGRATOR_FE_PAIR) {
    // two forward Euler steps: y -> y_mid (step starting at t) -> y_next
    // (step starting at t + d_t; its table slots follow three later)
    if (!use_fused || !p->fused[3] || !y_mid) return fail("no Euler pair kernel");
    rc = fused(3, y, y_mid, t, s_t, s_f, t_next, s_t + 3, s_t + 5);
  } else if (integrator == PML_INTEGRATOR_FORWARD_EULER) {
    rc = stage(0, y, nullptr, t, s_t, s_f);
  } else if (use_fused && integrator == PML_INTEGRATOR_EXPLICIT_MIDPOINT) {
    rc = fused(2, y, nullptr, t, s_t, s_h, t + half, s_h, s_f);
  } else if (use_fused) {
    // RK4: stages 1+2 write u_b (= u3) and acc; stages 3+4 write the slot
    rc = fused(0, y, ws->u_b, t, s_t, s_h, t + half, s_h, s_h);
    if (!rc) rc = fused(1, ws->u_b, nullptr, t + half, s_h, s_f, t + d_t, s_f, s_f);
  } else if (integrator == PML_INTEGRATOR_EXPLICIT_MIDPOINT) {
    rc = stage(1, y, ws->u_a, t, s_t, s_h);
    if (!rc) rc = stage(2, ws->u_a, nullptr, t + half, s_h, s_f);
  } else {
    rc = stage(3, y, ws->u_a, t, s_t, s_h);
    if (!rc) rc = stage(4, ws->u_a, ws->u_b, t + half, s_h, s_h);
    if (!rc) rc = stage(5, ws->u_b, ws->u_a, t + half, s_h, s_f);
    if (!rc) rc = stage(6, ws->u_a, nullptr, t + d_t, s_f, s_f);
  }
  return rc;
}

}  // namespace

extern "C" {

const char* pml_last_error(void) { return g_error.c_str(); }

int pml_version(void) { return 100; }

int pml_compile_to_cubin(const char* source, const char* cubin_path) {
  std::vector<char> cubin;
  std::string src_path = std::string(cubin_path) + ".cu";
  std::vector<char> text(source, source + std::strlen(source));
  write_file(src_path.c_str(), text);
  if (nvrtc_compile(source, cubin, src_path.c_str())) return -1;
  if (!write_file(cubin_path, cubin))
    return fail(std::string("cannot write ") + cubin_path);
  return 0;
}

int pml_plan_create(const char* source, const pml_plan_desc* desc,
                    const char* cubin_path, pml_plan** out) {
  if (!source || !desc || !out) return fail("null argument");
  if (load_driver()) return -1;
  std::vector<char> cubin;
  bool cached = cubin_path && read_file(cubin_path, cubin);
  std::string src_path = cubin_path ? std::string(cubin_path) + ".cu" : "";
  if (!cached) {
    if (cubin_path) {
      std::vector<char> text(source, source + std::strlen(source));
      write_file(src_path.c_str(), text);
    }
    if (nvrtc_compile(source, cubin,
                      cubin_path ? src_path.c_str() : "pml_generated.cu"))
      return -1;
    if (cubin_path) write_file(cubin_path, cubin);
  }
  pml_plan* p = new pml_plan();
  p->desc = *desc;
  CUresult r = g_drv.moduleLoadData(&p->module, cubin.data());
  if (r != CUDA_SUCCESS && cached) {
    // stale cache entry: recompile
    if (nvrtc_compile(source, cubin)) { delete p; return -1; }
    write_file(cubin_path, cubin);
    r = g_drv.moduleLoadData(&p->module, cubin.data());
  }
  if (r != CUDA_SUCCESS) {
    delete p;
    return fail("cuModuleLoadData: " + cu_err(r));
  }
  static const char* names[7] = {"pml_stage_fe",    "pml_stage_mid1",
                                 "pml_stage_mid2",  "pml_stage_rk4_1",
                                 "pml_stage_rk4_2", "pml_stage_rk4_3",
                                 "pml_stage_rk4_4"};
  for (int i = 0; i < 7; ++i) {
    r = g_drv.moduleGetFunction(&p->stage[i], p->module, names[i]);
    if (r != CUDA_SUCCESS) {
      std::string m = std::string("missing kernel ") + names[i];
      pml_plan_destroy(p);
      return fail(m);
    }
  }
  if (desc->fused) {
    static const char* fnames[3] = {"pml_fused_rk4_12", "pml_fused_rk4_34",
                                    "pml_fused_mid"};
    // rk4 1+2 and midpoint read only the input ring; rk4 3+4 adds the rings
    // of the step-start state and the accumulator
    p->fsmem[0] = (unsigned)desc->fused_smem[0];
    p->fsmem[1] = (unsigned)desc->fused_smem[1];
    p->fsmem[2] = (unsigned)desc->fused_smem[0];
    // two forward Euler steps per launch: column-marching variant only
    if (g_drv.moduleGetFunction(&p->fused[3], p->module, "pml_fused_fe2") ==
        CUDA_SUCCESS) {
      p->fsmem[3] = p->fsmem[0];
      CUresult q = CUDA_SUCCESS;
      if (p->fsmem[3] > 40 * 1024)
        q = g_drv.funcSetAttribute(p->fused[3],
                                   CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                   (int)p->fsmem[3]);
      if (q == CUDA_SUCCESS)
        q = g_drv.funcSetAttribute(
            p->fused[3], CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, 100);
      if (q != CUDA_SUCCESS) p->fused[3] = nullptr;
    } else {
      p->fused[3] = nullptr;
    }
    for (int i = 0; i < 3; ++i) {
      r = g_drv.moduleGetFunction(&p->fused[i], p->module, fnames[i]);
      if (r == CUDA_SUCCESS && p->fsmem[i] > 40 * 1024)  // static + dynamic > 48 KB needs the opt-in
        r = g_drv.funcSetAttribute(
            p->fused[i], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
            (int)p->fsmem[i]);
      if (r == CUDA_SUCCESS)
        r = g_drv.funcSetAttribute(
            p->fused[i], CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT,
            100);
      if (r != CUDA_SUCCESS) {
        std::string m = std::string("fused kernel unavailable: ") + fnames[i] +
                        ": " + cu_err(r);
        pml_plan_destroy(p);
        return fail(m);
      }
    }
    auto cdivf = [](int a, int d) { return (unsigned)((a + d - 1) / d); };
    p->fblock = dim3((unsigned)desc->fused_threads, 1, 1);
    const int* n = desc->shape;
    const int ftx = desc->fused_tile[0], fty = desc->fused_tile[1];
    if (desc->n_dims == 3)
      p->fgrid = dim3(cdivf(n[2], ftx), cdivf(n[1], fty),
                      cdivf(n[0], desc->fused_zc));
    else
      p->fgrid = dim3(cdivf(n[1], ftx), cdivf(n[0], desc->fused_zc), 1);
  }
  if (desc->small_threads > 0) {
    r = g_drv.moduleGetFunction(&p->small_run, p->module, "pml_small_run");
    if (r != CUDA_SUCCESS) {
      pml_plan_destroy(p);
      return fail("missing kernel pml_small_run");
    }
  }
  r = g_drv.moduleGetFunction(&p->eval_rhs, p->module, "pml_eval_rhs");
  if (r != CUDA_SUCCESS) {
    pml_plan_destroy(p);
    return fail("missing kernel pml_eval_rhs");
  }
  // (cubins of older template versions may lack it: optional)
  if (g_drv.moduleGetFunction(&p->apply_dir, p->module, "pml_apply_dirichlet_planes") !=
      CUDA_SUCCESS)
    p->apply_dir = nullptr;
  if (desc->n_lap > 0) {
    struct { const char* n; CUfunction* f; } js[] = {
        {"pml_jacobi_init", &p->jac_init}, {"pml_jacobi_sweep", &p->jac_sweep},
        {"pml_jacobi_store", &p->jac_store}, {"pml_jacobi_loop", &p->jac_loop}};
    for (auto& j : js) {
      r = g_drv.moduleGetFunction(j.f, p->module, j.n);
      if (r != CUDA_SUCCESS) {
        std::string m = std::string("missing kernel ") + j.n;
        pml_plan_destroy(p);
        return fail(m);
      }
    }
  }
  const int* n = desc->shape;
  const int* b = desc->block;
  p->block = dim3(b[0], b[1], b[2]);
  auto cdiv = [](int a, int d) { return (unsigned)((a + d - 1) / d); };
  if (desc->n_dims <= 1)
    p->grid = dim3(cdiv(n[0], b[0]), 1, 1);
  else if (desc->n_dims == 2)
    p->grid = dim3(cdiv(n[1], b[0]), cdiv(n[0], b[1]), 1);
  else
    p->grid = dim3(cdiv(n[2], b[0]), cdiv(n[1], b[1]), cdiv(n[0], b[2]));
  p->sgrid = p->grid;
  const int zrep = desc->zrep > 1 ? desc->zrep : 1;
  if (desc->n_dims == 2)
    p->sgrid.y = cdiv(n[0], b[1] * zrep);
  else if (desc->n_dims == 3)
    p->sgrid.z = cdiv(n[0], b[2] * zrep);
  p->jgrid = p->grid;
  if (desc->n_dims == 2)
    p->jgrid.y = cdiv(n[0], b[1] * PML_JACOBI_REP);
  else if (desc->n_dims == 3)
    p->jgrid.z = cdiv(n[0], b[2] * PML_JACOBI_REP);
  p->n_cells = (long long)n[0] * n[1] * n[2];
  p->n_blocks = (long long)p->grid.x * p->grid.y * p->grid.z;
  *out = p;
  return 0;
}

int pml_plan_destroy(pml_plan* p) {
  if (!p) return 0;
  if (p->module && g_drv.moduleUnload) g_drv.moduleUnload(p->module);
  delete p;
  return 0;
}

int pml_plan_set_tables(pml_plan* p, const pml_tables* t) {
  if (!p || !t) return fail("null argument");
  p->tables = *t;
  return 0;
}

long long pml_plan_launches(const pml_plan* p) { return p ? p->launches : 0; }

static int fdm_run_batch(pml_plan* p, int integrator, const pml_workspace* ws,
                         const double* y0, double* traj, long long stride,
                         int batch, long long y_batch_stride,
                         long long traj_batch_stride, long long ws_batch_stride,
                         const double* t_host, int n_steps, double d_t,
                         long long slot0, const double* jacobi_init,
                         double jacobi_tol, long long max_sweeps,
                         int* sweeps_out, void* stream);

int pml_fdm_run(pml_plan* p, int integrator, const pml_workspace* ws,
                const double* y0, double* traj, long long stride,
                const double* t_host, int n_steps, double d_t, long long slot0,
                const double* jacobi_init, double jacobi_tol,
                long long max_sweeps, int* sweeps_out, void* stream) {
  return fdm_run_batch(p, integrator, ws, y0, traj, stride, 1, 0, 0, 0, t_host,
                       n_steps, d_t, slot0, jacobi_init, jacobi_tol, max_sweeps,
                       sweeps_out, stream);
}

int pml_fdm_run_batch(pml_plan* p, int integrator, const pml_workspace* ws,
                      const double* y0, double* traj, long long stride,
                      int batch, long long y_batch_stride,
                      long long traj_batch_stride, long long ws_batch_stride,
                      const double* t_host, int n_steps, double d_t,
                      long long slot0, void* stream) {
  if (batch < 1) return fail("empty batch");
  if (p && p->desc.n_lap > 0)
    return fail("batched solves of systems with Y_LAPLACIAN equations");
  return fdm_run_batch(p, integrator, ws, y0, traj, stride, batch,
                       y_batch_stride, traj_batch_stride, ws_batch_stride, t_host,
                       n_steps, d_t, slot0, nullptr, 0.0, 0, nullptr, stream);
}

static int fdm_run_batch(pml_plan* p, int integrator, const pml_workspace* ws,
                         const double* y0, double* traj, long long stride,
                         int batch, long long y_batch_stride,
                         long long traj_batch_stride, long long ws_batch_stride,
                         const double* t_host, int n_steps, double d_t,
                         long long slot0, const double* jacobi_init,
                         double jacobi_tol, long long max_sweeps,
                         int* sweeps_out, void* stream) {
  if (!p || !ws || !y0 || !traj || !t_host) return fail("null argument");
  if (integrator < 0 || integrator > 2) return fail("unknown integrator");
  if (p->desc.n_lap > 0 && !jacobi_init)
    return fail("Y_LAPLACIAN equations need a Jacobi start array");
  CUstream s = (CUstream)stream;
  PmlArgs a;
  std::memset(&a, 0, sizeof(a));
  fill_tables(p, a);
  a.dt = d_t;
  a.lap_rhs = ws->lap_rhs;
  const long long lap_elems = (long long)p->desc.n_lap * p->n_cells;
  if (p->small_run && p->desc.n_lap == 0 && ws->t_dev && ws->t_capacity > 0) {
    // small mesh: all steps of a chunk in one single-block launch
    for (int first = 0; first < n_steps; first += (int)ws->t_capacity) {
      const int count = n_steps - first < ws->t_capacity
                            ? n_steps - first : (int)ws->t_capacity;
      PML_CUDA(cudaMemcpyAsync(ws->t_dev, t_host + first, count * sizeof(double),
                               cudaMemcpyHostToDevice, (cudaStream_t)s));
      PmlSmallArgs f;
      f.s = a;
      for (int q = 0; q < 6; ++q) {
        f.neu_stride[q] = p->tables.neu_stride[q];
        f.dir_stride[q] = p->tables.dir_stride[q];
      }
      f.s.y = first == 0 ? y0 : traj + (long long)(first - 1) * stride;
      f.traj = traj + (long long)first * stride;
      f.stride = stride;
      f.t = ws->t_dev;
      f.n_steps = count;
      f.integrator = integrator;
      f.slot0 = slot0 + 3LL * first;
      f.u_a = ws->u_a;
      f.u_b = ws->u_b;
      f.acc = ws->acc;
      f.y_batch_stride = first == 0 ? y_batch_stride : traj_batch_stride;
      f.traj_batch_stride = traj_batch_stride;
      f.ws_batch_stride = ws_batch_stride;
      void* sparams[] = {&f};
      PML_CU(g_drv.launchKernel(p->small_run, (unsigned)batch, 1, 1,
                                (unsigned)p->desc.small_threads, 1, 1, 0, s,
                                sparams, nullptr));
      p->launches += 1;
      // t_host may be reused by the caller and t_dev by the next chunk
      PML_CUDA(cudaStreamSynchronize((cudaStream_t)s));
    }
    return 0;
  }
  // large meshes: every member saturates the device, members run in turn
  // (they share the scratch buffers)
  for (int b = 1; b < batch; ++b) {
    if (fdm_run_batch(p, integrator, ws, y0 + b * y_batch_stride,
                      traj + b * traj_batch_stride, stride, 1, 0, 0, 0, t_host,
                      n_steps, d_t, slot0, nullptr, 0.0, 0, nullptr, stream))
      return -1;
  }
  auto aligned16 = [](const void* q) { return ((uintptr_t)q & 15u) == 0; };
  for (int j = 0; j < n_steps; ++j) {
    const double t = t_host[j];
    const double* y = j == 0 ? y0 : traj + (long long)(j - 1) * stride;
    double* y_next = traj + (long long)j * stride;
    const long long s_t = slot0 + 3LL * j, s_f = s_t + 2;
    if (integrator == PML_INTEGRATOR_FORWARD_EULER && p->desc.fused &&
        p->fused[3] && j + 1 < n_steps && p->desc.n_alg == 0 &&
        p->desc.n_lap == 0 &&
        aligned16(y) && aligned16(y_next) && aligned16(y_next + stride) &&
        aligned16(ws->u_b) && aligned16(ws->acc)) {
      // steps j and j + 1 in one launch of the stage-pair kernel
      if (run_step(p, PML_INTEGRATOR_FE_PAIR, ws, a, y, y_next + stride, t, d_t,
                   s_t, -1, nullptr, s, 0, -1, y_next, t_host[j + 1]))
        return -1;
      ++j;
      continue;
    }
    if (run_step(p, integrator, ws, a, y, y_next, t, d_t, s_t, -1, nullptr, s))
      return -1;
    if (p->desc.n_lap > 0) {
      PmlArgs tbl = a;
      bind_neu(p, tbl.neu, s_f);
      bind_dir(p, tbl.dir, s_f);
      int sweeps = 0;
      if (jacobi_solve(p, ws, tbl, ws->lap_rhs, jacobi_init + (long long)j * lap_elems,
                       y_next, jacobi_tol, max_sweeps, &sweeps, s))
        return -1;
      if (sweeps_out) sweeps_out[j] = sweeps;
    }
  }
  return 0;
}

int pml_fdm_phase_count(const pml_plan* p, int integrator) {
  if (!p) return fail("null argument");
  if (integrator == PML_INTEGRATOR_FORWARD_EULER) return 1;
  if (integrator == PML_INTEGRATOR_EXPLICIT_MIDPOINT) return p->desc.fused ? 1 : 2;
  if (integrator == PML_INTEGRATOR_RK4) return p->desc.fused ? 2 : 4;
  return fail("unknown integrator");
}

int pml_fdm_phase(pml_plan* p, int integrator, const pml_workspace* ws,
                  const double* y, double* y_next, double t, double d_t,
                  long long slot0, int phase, double** fresh_out, void* stream) {
  return pml_fdm_phase_planes(p, integrator, ws, y, y_next, t, d_t, slot0, phase,
                              0, -1, fresh_out, stream);
}

int pml_fdm_phase_planes(pml_plan* p, int integrator, const pml_workspace* ws,
                         const double* y, double* y_next, double t, double d_t,
                         long long slot0, int phase, int z_begin, int z_end,
                         double** fresh_out, void* stream) {
  if (!p || !ws || !y || !y_next) return fail("null argument");
  if (integrator < 0 || integrator > 2) return fail("unknown integrator");
  if (p->desc.n_lap > 0 || p->desc.n_alg > 0)
    return fail("phase-wise stepping needs a fully time-stepped system");
  if (phase < 0 || phase >= pml_fdm_phase_count(p, integrator))
    return fail("phase out of range");
  PmlArgs a;
  std::memset(&a, 0, sizeof(a));
  fill_tables(p, a);
  a.dt = d_t;
  a.lap_rhs = ws->lap_rhs;
  return run_step(p, integrator, ws, a, y, y_next, t, d_t, slot0, phase,
                  fresh_out, (CUstream)stream, z_begin, z_end);
}

int pml_eval_rhs(pml_plan* p, const double* u, double* out, double t,
                 long long slot, void* stream) {
  if (!p || !u || !out) return fail("null argument");
  PmlArgs a;
  std::memset(&a, 0, sizeof(a));
  fill_tables(p, a);
  a.u = u;
  a.y = u;
  a.u_out = out;
  a.t_eval = t;
  bind_neu(p, a.neu, slot);
  bind_dir(p, a.dir, slot);
  bind_dir(p, a.dir_full, slot);
  void* params[] = {&a};
  return launch(p, p->eval_rhs, params, (CUstream)stream);
}

int pml_apply_dirichlet(pml_plan* p, double* planes, long long slot, void* stream) {
  if (!p || !planes) return fail("null argument");
  if (!p->apply_dir) return fail("plan has no Dirichlet kernel");
  PmlArgs a;
  std::memset(&a, 0, sizeof(a));
  fill_tables(p, a);
  bind_dir(p, a.dir, slot);
  a.u_out = planes;
  void* params[] = {&a};
  return launch(p, p->apply_dir, params, (CUstream)stream);
}

int pml_jacobi_run(pml_plan* p, const pml_workspace* ws, const double* rhs,
                   const double* y_init, double* y_next, long long slot,
                   double tol, long long max_sweeps, int* sweeps_out,
                   void* stream) {
  if (!p || !ws || !rhs || !y_init || !y_next) return fail("null argument");
  if (p->desc.n_lap <= 0) return fail("plan has no Y_LAPLACIAN equations");
  PmlArgs a;
  std::memset(&a, 0, sizeof(a));
  fill_tables(p, a);
  bind_neu(p, a.neu, slot);
  bind_dir(p, a.dir, slot);
  bind_dir(p, a.dir_full, slot);
  return jacobi_solve(p, ws, a, rhs, y_init, y_next, tol, max_sweeps,
                      sweeps_out, (CUstream)stream);
}

}  // extern "C"

// ===========================================================================
// fixed sm_100a kernels
// ===========================================================================
namespace {

constexpr int kThreads = 256;

// (state, cell, comp) channels-last  <->  (state, comp, cell) planes.  Reads and
// writes are both coalesced: consecutive threads walk the contiguous side and
// the strided side stays within y_dim * 8 bytes of each other (L1/L2 merge).
__global__ void __launch_bounds__(kThreads)
aos_to_soa_kernel(const double* __restrict__ aos, double* __restrict__ soa,
                  long long n_cells, int c_dim, long long total) {
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < total;
       i += (long long)gridDim.x * kThreads) {
    const long long per_state = n_cells * c_dim;
    const long long st = i / per_state, r = i - st * per_state;
    const long long c = r / n_cells, cell = r - c * n_cells;
    soa[i] = __ldg(aos + st * per_state + cell * c_dim + c);
  }
}

__global__ void __launch_bounds__(kThreads)
soa_to_aos_kernel(const double* __restrict__ soa, double* __restrict__ aos,
                  long long n_cells, int c_dim, long long total) {
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < total;
       i += (long long)gridDim.x * kThreads) {
    const long long per_state = n_cells * c_dim;
    const long long st = i / per_state, r = i - st * per_state;
    const long long cell = r / c_dim, c = r - cell * c_dim;
    aos[i] = __ldg(soa + st * per_state + c * n_cells + cell);
  }
}

__global__ void __launch_bounds__(kThreads)
correction_kernel(const double* __restrict__ f, const double* __restrict__ g,
                  double* __restrict__ corr, long long n) {
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kThreads)
    corr[i] = __ldg(f + i) - __ldg(g + i);
}

// new = g + corr and per-block partial sums of (new - old)^2 per component
__global__ void __launch_bounds__(kThreads)
update_kernel(const double* __restrict__ g, const double* __restrict__ corr,
              const double* __restrict__ old_end, double* __restrict__ new_end,
              double* __restrict__ partials, long long n_cells) {
  const int c = blockIdx.y;
  const long long base = (long long)c * n_cells;
  double sq = 0.0;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n_cells;
       i += (long long)gridDim.x * kThreads) {
    const double v = __ldg(g + base + i) + __ldg(corr + base + i);
    const double d = v - __ldg(old_end + base + i);
    new_end[base + i] = v;
    sq += d * d;
  }
  __shared__ double red[kThreads / 32];
  for (int off = 16; off > 0; off >>= 1) sq += __shfl_down_sync(0xffffffffu, sq, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) s += red[w];
    partials[(long long)c * gridDim.x + blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(kThreads)
finish_sumsq_kernel(const double* __restrict__ partials, int n_partials,
                    double* __restrict__ sumsq) {
  const int c = blockIdx.x;
  __shared__ double red[kThreads];
  double s = 0.0;
  for (int i = threadIdx.x; i < n_partials; i += kThreads)
    s += partials[(long long)c * n_partials + i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int off = kThreads / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) sumsq[c] = red[0];
}

__global__ void __launch_bounds__(kThreads)
delta_kernel(const double* __restrict__ new_end, const double* __restrict__ last,
             double* __restrict__ delta, long long n) {
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kThreads)
    delta[i] = __ldg(new_end + i) - __ldg(last + i);
}

__global__ void __launch_bounds__(kThreads)
shift_kernel(double* __restrict__ traj, long long n_steps, long long stride,
             const double* __restrict__ delta, long long n) {
  const long long step = blockIdx.y;
  double* row = traj + step * stride;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n;
       i += (long long)gridDim.x * kThreads)
    row[i] += __ldg(delta + i);
  (void)n_steps;
}

struct IcMesh {
  int n_dims;
  int shape[3];
  int coord;
  const double* axis[3];
  const double* trig[4];
};
struct IcGaussian {
  pml_ic_gaussian_params comp[PML_IC_MAX_COMPONENTS];
};
struct IcFactors {
  const double* f[PML_IC_MAX_COMPONENTS * 3];
  double multiplier[PML_IC_MAX_COMPONENTS];
};

__device__ __forceinline__ void ic_cell(const IcMesh& m, long long cell, int& i0,
                                        int& i1, int& i2) {
  const int n1 = m.shape[1], n2 = m.shape[2];
  i2 = (int)(cell % n2);
  const long long r = cell / n2;
  i1 = (int)(r % n1);
  i0 = (int)(r / n1);
}

// Cartesian coordinates of a vertex, formed like mesh.py:to_cartesian_coordinates
// (trigonometric factors evaluated by the host, products in the same order)
__device__ __forceinline__ void ic_cartesian(const IcMesh& m, int i0, int i1,
                                             int i2, double* x) {
  const double a0 = __ldg(m.axis[0] + i0);
  const double a1 = m.n_dims > 1 ? __ldg(m.axis[1] + i1) : 0.0;
  const double a2 = m.n_dims > 2 ? __ldg(m.axis[2] + i2) : 0.0;
  if (m.coord == 0) {
    x[0] = a0;
    x[1] = a1;
    x[2] = a2;
  } else if (m.coord == 3) {
    const double ct = __ldg(m.trig[0] + i1), st = __ldg(m.trig[1] + i1);
    const double sp = __ldg(m.trig[2] + i2), cp = __ldg(m.trig[3] + i2);
    x[0] = a0 * sp * ct;
    x[1] = a0 * sp * st;
    x[2] = a0 * cp;
  } else {
    const double ct = __ldg(m.trig[0] + i1), st = __ldg(m.trig[1] + i1);
    x[0] = a0 * ct;
    x[1] = a0 * st;
    x[2] = a2;
  }
}

__global__ void __launch_bounds__(kThreads)
ic_gaussian_kernel(const __grid_constant__ IcMesh m, int y_dim,
                   const __grid_constant__ IcGaussian g, double* __restrict__ planes,
                   long long n_cells) {
  for (long long cell = blockIdx.x * (long long)kThreads + threadIdx.x;
       cell < n_cells; cell += (long long)gridDim.x * kThreads) {
    int i0, i1, i2;
    ic_cell(m, cell, i0, i1, i2);
    double x[3];
    ic_cartesian(m, i0, i1, i2, x);
    const int d = m.n_dims;
    for (int c = 0; c < y_dim; ++c) {
      const pml_ic_gaussian_params& p = g.comp[c];
      double maha = 0.0;
      for (int k = 0; k < d; ++k) {
        double w = 0.0;
        for (int j = 0; j < d; ++j) w += (x[j] - p.mean[j]) * p.whiten[j * d + k];
        maha += w * w;
      }
      planes[(long long)c * n_cells + cell] =
          exp(-0.5 * (p.log_norm + maha)) * p.multiplier;
    }
  }
}

__global__ void __launch_bounds__(kThreads)
ic_separable_kernel(const __grid_constant__ IcMesh m, int y_dim,
                    const __grid_constant__ IcFactors f, double* __restrict__ planes,
                    long long n_cells) {
  for (long long cell = blockIdx.x * (long long)kThreads + threadIdx.x;
       cell < n_cells; cell += (long long)gridDim.x * kThreads) {
    int i[3];
    ic_cell(m, cell, i[0], i[1], i[2]);
    for (int c = 0; c < y_dim; ++c) {
      double v = __ldg(f.f[c * m.n_dims] + i[0]);
      for (int a = 1; a < m.n_dims; ++a) v *= __ldg(f.f[c * m.n_dims + a] + i[a]);
      planes[(long long)c * n_cells + cell] = v * f.multiplier[c];
    }
  }
}

int ic_mesh_from(const pml_ic_mesh* mesh, IcMesh& m, long long& n_cells) {
  if (!mesh || mesh->n_dims < 1 || mesh->n_dims > 3)
    return fail("initial conditions on the device need a mesh with 1..3 axes");
  m.n_dims = mesh->n_dims;
  m.coord = mesh->coord;
  n_cells = 1;
  for (int a = 0; a < 3; ++a) {
    m.shape[a] = a < mesh->n_dims ? mesh->shape[a] : 1;
    m.axis[a] = mesh->axis_dev[a];
    n_cells *= m.shape[a];
  }
  for (int a = 0; a < 4; ++a) m.trig[a] = mesh->trig_dev[a];
  return 0;
}

unsigned grid_for(long long n, unsigned cap = 148 * 16) {
  long long b = (n + kThreads - 1) / kThreads;
  if (b < 1) b = 1;
  return (unsigned)(b < cap ? b : cap);
}

}  // namespace

extern "C" {

int pml_aos_to_soa(const double* aos, double* soa, long long n_cells, int y_dim,
                   long long n_states, void* stream) {
  const long long total = n_cells * y_dim * n_states;
  if (total == 0) return 0;
  aos_to_soa_kernel<<<grid_for(total), kThreads, 0, (cudaStream_t)stream>>>(
      aos, soa, n_cells, y_dim, total);
  PML_CUDA(cudaGetLastError());
  return 0;
}

int pml_soa_to_aos(const double* soa, double* aos, long long n_cells, int y_dim,
                   long long n_states, void* stream) {
  const long long total = n_cells * y_dim * n_states;
  if (total == 0) return 0;
  soa_to_aos_kernel<<<grid_for(total), kThreads, 0, (cudaStream_t)stream>>>(
      soa, aos, n_cells, y_dim, total);
  PML_CUDA(cudaGetLastError());
  return 0;
}

int pml_ic_gaussian(const pml_ic_mesh* mesh, int y_dim,
                    const pml_ic_gaussian_params* params, double* planes,
                    void* stream) {
  if (!params || !planes) return fail("null argument");
  if (y_dim < 1 || y_dim > PML_IC_MAX_COMPONENTS)
    return fail("too many components for a device-side initial condition");
  IcMesh m;
  long long n_cells;
  if (ic_mesh_from(mesh, m, n_cells)) return -1;
  IcGaussian g;
  std::memset(&g, 0, sizeof(g));
  for (int c = 0; c < y_dim; ++c) g.comp[c] = params[c];
  ic_gaussian_kernel<<<grid_for(n_cells), kThreads, 0, (cudaStream_t)stream>>>(
      m, y_dim, g, planes, n_cells);
  PML_CUDA(cudaGetLastError());
  return 0;
}

int pml_ic_separable(const pml_ic_mesh* mesh, int y_dim,
                     const double* const* factors, const double* multipliers,
                     double* planes, void* stream) {
  if (!factors || !multipliers || !planes) return fail("null argument");
  if (y_dim < 1 || y_dim > PML_IC_MAX_COMPONENTS)
    return fail("too many components for a device-side initial condition");
  IcMesh m;
  long long n_cells;
  if (ic_mesh_from(mesh, m, n_cells)) return -1;
  if (m.coord != 0) return fail("separable initial conditions need a Cartesian mesh");
  IcFactors f;
  std::memset(&f, 0, sizeof(f));
  for (int c = 0; c < y_dim; ++c) {
    f.multiplier[c] = multipliers[c];
    for (int a = 0; a < m.n_dims; ++a) f.f[c * m.n_dims + a] = factors[c * m.n_dims + a];
  }
  ic_separable_kernel<<<grid_for(n_cells), kThreads, 0, (cudaStream_t)stream>>>(
      m, y_dim, f, planes, n_cells);
  PML_CUDA(cudaGetLastError());
  return 0;
}

int pml_parareal_correction(const double* f, const double* g, double* corr,
                            long long n, void* stream) {
  correction_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(f, g, corr, n);
  PML_CUDA(cudaGetLastError());
  return 0;
}

int pml_parareal_update(const double* g, const double* corr, const double* old_end,
                        double* new_end, double* sumsq, double* scratch,
                        long long n_cells, int y_dim, void* stream) {
  unsigned blocks = grid_for(n_cells, 1024);
  dim3 grid(blocks, (unsigned)y_dim);
  update_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(g, corr, old_end, new_end,
                                                             scratch, n_cells);
  PML_CUDA(cudaGetLastError());
  finish_sumsq_kernel<<<(unsigned)y_dim, kThreads, 0, (cudaStream_t)stream>>>(
      scratch, (int)blocks, sumsq);
  PML_CUDA(cudaGetLastError());
  return 0;
}

int pml_parareal_shift(double* traj, long long n_steps, long long stride,
                       const double* new_end, double* delta, long long n,
                       void* stream) {
  if (n_steps <= 0) return 0;
  delta_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(
      new_end, traj + (n_steps - 1) * stride, delta, n);
  PML_CUDA(cudaGetLastError());
  for (long long first = 0; first < n_steps; first += 32768) {
    const long long count = n_steps - first < 32768 ? n_steps - first : 32768;
    dim3 grid(grid_for(n, 148 * 4), (unsigned)count);
    shift_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(
        traj + first * stride, count, stride, delta, n);
    PML_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // extern "C"
