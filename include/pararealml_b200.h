/*
 * C ABI of the B200-native FDM + Parareal hot path (libpararealml_b200.so).
 *
 * The reference (ViktorC/PararealML v0.3.0) is pure Python and has no FFI; the
 * entry points below are what a binding for its hot path would call, each
 * citing the reference code it replaces.  All pointers named *_dev are device
 * pointers owned by the caller (PyTorch tensors in the shipped host layer);
 * nothing is allocated per call.  Functions return 0 on success and a negative
 * code on failure, pml_last_error() then holds a message.  A plan is
 * thread-compatible, not thread-safe.
 */
#ifndef PARAREALML_B200_H
#define PARAREALML_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pml_plan pml_plan;

enum { PML_INTEGRATOR_FORWARD_EULER = 0, PML_INTEGRATOR_EXPLICIT_MIDPOINT = 1,
       PML_INTEGRATOR_RK4 = 2 };

typedef struct pml_plan_desc {
  int n_dims;       /* 0 (ODE) .. 3 */
  int shape[3];     /* mesh vertices per axis, trailing unused axes = 1 */
  int y_dim;        /* components of y */
  int n_dt, n_alg, n_lap; /* equations per LHS kind (differential_equation.py:140-149) */
  int block[3];     /* thread block shape the source was generated for */
  int fused;        /* != 0: the source has the fused stage-pair kernels
                       (1 = one thread per cell of the stage-A tile, 2 = column-
                       marching threads with register columns; the latter also
                       has the forward Euler step-pair kernel) */
  int fused_tile[2]; /* their tile (cells along the contiguous axis, axis 1) */
  int fused_zc;     /* planes of the marching axis per thread block */
  int fused_threads; /* threads per block of the fused kernels */
  int fused_smem[2]; /* dynamic shared memory: stage 1+2 / midpoint, stage 3+4 */
  int small_threads; /* > 0: the source has the single-block time loop kernel */
  int zrep;         /* cells along axis 0 per thread in the stage kernels */
} pml_plan_desc;

/* NaN-coded boundary tables and 1-D coordinate vectors (device pointers).
 * Face index = axis * 2 + side.  Replaces the Constraint objects of
 * constrained_problem.py:303-476 / constraint.py:6-101 on the device. */
typedef struct pml_tables {
  const double* neu[6];
  long long neu_stride[6]; /* doubles between time slots, 0 = static */
  const double* dir[6];
  long long dir_stride[6];
  const double* coord[3];  /* mesh.vertex_axis_coordinates (mesh.py:353) */
  const double* aux[4];    /* 1/r[i0], sin(phi)[i2], cos(phi)[i2], 1/sin(phi)[i2] */
} pml_tables;

/* Scratch buffers, each y_dim * n_cells doubles unless noted. */
typedef struct pml_workspace {
  double* u_a;
  double* u_b;
  double* acc;
  double* lap_rhs;   /* n_lap * n_cells */
  double* jac_a;     /* n_lap * n_cells */
  double* jac_b;     /* n_lap * n_cells */
  double* partials;  /* max(2, one double per thread block of the stage grid) */
  int* flags;        /* 4 ints: done, sweeps, block ticket of the sweep kernel,
                        barrier counter of the persistent Jacobi loop */
  double* t_dev;     /* step start times for the single-block kernel */
  long long t_capacity; /* doubles available at t_dev */
} pml_workspace;

const char* pml_last_error(void);
int pml_version(void);

/* Compiles (NVRTC, sm_100a) the generated stage-kernel source, or loads the
 * cubin cached at cubin_path if that file exists (it is written otherwise;
 * may be NULL).  Replaces FDMSymbolMapper construction + sp.lambdify
 * (symbol_mapper.py:28-42, 222-253). */
int pml_plan_create(const char* source, const pml_plan_desc* desc,
                    const char* cubin_path, pml_plan** plan);
/* NVRTC only, no GPU needed: compiles to a cubin file (build-time check). */
int pml_compile_to_cubin(const char* source, const char* cubin_path);
int pml_plan_destroy(pml_plan* plan);
int pml_plan_set_tables(pml_plan* plan, const pml_tables* tables);
long long pml_plan_launches(const pml_plan* plan); /* kernels launched so far */

/* n_steps explicit time steps, asynchronous on `stream` unless n_lap > 0.
 * Step j reads y0_dev (j == 0) or trajectory slot j-1 and writes slot j.
 * t_host[j] is the step's start time (operator.py:60-74); boundary table slot
 * of step j, stage time q in {t, t + dt/2, t + dt} is slot0 + 3 j + q.
 * Replaces the time loop and y_next function of fdm_operator.py:48-165 and
 * the integrators of numerical_integrator.py:47-132. */
int pml_fdm_run(pml_plan* plan, int integrator, const pml_workspace* ws,
                const double* y0_dev, double* traj_dev,
                long long traj_step_stride, const double* t_host, int n_steps,
                double d_t, long long slot0, const double* jacobi_init_dev,
                double jacobi_tol, long long max_sweeps, int* sweeps_out_host,
                void* stream);

/* The same for `batch` initial states of one problem over one time grid
 * (the data generation of supervised_ml_operator.py:130-236 solves many
 * perturbed sub-IVPs of a problem): member b reads y0_dev + b *
 * y0_batch_stride and writes traj_dev + b * traj_batch_stride.  Small meshes
 * run all members in ONE launch, one thread block per member, and need
 * `batch` sets of scratch buffers (u_a, u_b, acc at member stride
 * ws_batch_stride doubles); larger meshes run member after member.  Systems
 * with Y_LAPLACIAN equations are not batched. */
int pml_fdm_run_batch(pml_plan* plan, int integrator, const pml_workspace* ws,
                      const double* y0_dev, double* traj_dev,
                      long long traj_step_stride, int batch,
                      long long y0_batch_stride, long long traj_batch_stride,
                      long long ws_batch_stride, const double* t_host,
                      int n_steps, double d_t, long long slot0, void* stream);

/* The same time step one launch ("phase") at a time, for callers that
 * decompose the mesh into slabs of axis 0 and exchange halo planes between
 * launches (fully time-stepped systems only).  pml_fdm_phase_count: launches
 * per step (RK4: 2 stage-pair launches, or 4 stage launches).  pml_fdm_phase
 * runs phase `phase` of the step starting at t; *fresh_out_dev receives the
 * device buffer that phase has written -- the next phase's stencil input
 * (workspace memory) or y_next_dev after the last phase. */
int pml_fdm_phase_count(const pml_plan* plan, int integrator);
int pml_fdm_phase(pml_plan* plan, int integrator, const pml_workspace* ws,
                  const double* y_dev, double* y_next_dev, double t, double d_t,
                  long long slot0, int phase, double** fresh_out_dev,
                  void* stream);
/* The same for the planes [z_begin, z_end) of axis 0 only (stage-pair kernels;
 * z_end < 0: up to the last plane): a slab-decomposed caller launches the
 * planes its neighbours are waiting for first and exchanges them while the
 * launch of the remaining planes runs.  The launches of one phase may be issued
 * in any order; together they must cover every plane exactly once. */
int pml_fdm_phase_planes(pml_plan* plan, int integrator, const pml_workspace* ws,
                         const double* y_dev, double* y_next_dev, double t,
                         double d_t, long long slot0, int phase, int z_begin,
                         int z_end, double** fresh_out_dev, void* stream);

/* One evaluation of the generated right-hand sides (differentiator entry
 * points, numerical_differentiator.py:114-870). */
int pml_eval_rhs(pml_plan* plan, const double* u_dev, double* out_dev, double t,
                 long long slot, void* stream);

/* Jacobi anti-Laplacian on its own (numerical_differentiator.py:872-927). */
int pml_jacobi_run(pml_plan* plan, const pml_workspace* ws,
                   const double* rhs_dev, const double* y_init_dev,
                   double* y_next_dev, long long slot, double tol,
                   long long max_sweeps, int* sweeps_out_host, void* stream);

/* Layout conversion between the reference's channels-last (*mesh, C) arrays
 * and the device's component planes, batched over n_states. */
int pml_aos_to_soa(const double* aos_dev, double* soa_dev, long long n_cells,
                   int y_dim, long long n_states, void* stream);
int pml_soa_to_aos(const double* soa_dev, double* aos_dev, long long n_cells,
                   int y_dim, long long n_states, void* stream);

/* Parareal device helpers (parareal_operator.py:164, 183-185, 85-100, 192).
 * States are component planes of n_cells doubles. */
int pml_parareal_correction(const double* fine_end_dev,
                            const double* coarse_end_dev, double* corr_dev,
                            long long n, void* stream);
/* new_end = coarse_end + corr; sumsq_dev[c] = sum over cells of
 * (new_end - old_end)^2 for component c (deterministic two-pass reduction;
 * scratch_dev needs y_dim * 1024 doubles). */
int pml_parareal_update(const double* coarse_end_dev, const double* corr_dev,
                        const double* old_end_dev, double* new_end_dev,
                        double* sumsq_dev, double* scratch_dev,
                        long long n_cells, int y_dim, void* stream);
/* traj[s] += new_end - traj[n_steps - 1] for every step s (rigid shift). */
int pml_parareal_shift(double* traj_dev, long long n_steps,
                       long long step_stride, const double* new_end_dev,
                       double* delta_scratch_dev, long long n, void* stream);

/* Device-side initial conditions (initial_condition.py:246-378): the state
 * at t0 is evaluated on the mesh straight into component planes, so a 512^3
 * Gaussian costs a millisecond instead of minutes of host NumPy.
 * axis_dev[a]: mesh.vertex_axis_coordinates[a]; for curvilinear meshes
 * trig_dev holds the host-evaluated cos / sin of the angular axes
 * (cos(theta)[i1], sin(theta)[i1], sin(phi)[i2], cos(phi)[i2]) so that the
 * Cartesian coordinates are the products mesh.py:to_cartesian_coordinates
 * forms. */
typedef struct pml_ic_mesh {
  int n_dims;
  int shape[3];
  int coord;                  /* 0 cartesian, 1 polar, 2 cylindrical, 3 spherical */
  const double* axis_dev[3];
  const double* trig_dev[4];
} pml_ic_mesh;

#define PML_IC_MAX_COMPONENTS 16
/* One multivariate normal density per component (GaussianInitialCondition,
 * initial_condition.py:300-343; scipy.stats.multivariate_normal.pdf):
 * exp(-0.5 (log_norm + |(x - mean) U|^2)) * multiplier with the whitening
 * matrix U (row-major x_dim x x_dim) and log_norm = rank log(2 pi) +
 * log pdet(cov), both computed by the host from the covariance like SciPy. */
typedef struct pml_ic_gaussian_params {
  double mean[3];
  double whiten[9];
  double log_norm;
  double multiplier;
} pml_ic_gaussian_params;
int pml_ic_gaussian(const pml_ic_mesh* mesh, int y_dim,
                    const pml_ic_gaussian_params* params_host,
                    double* planes_dev, void* stream);
/* Product over the mesh axes of per-axis factor vectors (Cartesian meshes:
 * MarginalBetaProductInitialCondition, initial_condition.py:346-378, with the
 * 1-D Beta densities evaluated by the host): planes[c][cell] =
 * ((f[c][0][i0] * f[c][1][i1]) * f[c][2][i2]) * multiplier[c];
 * factors_dev[c * n_dims + a] is a device vector of shape[a] doubles. */
int pml_ic_separable(const pml_ic_mesh* mesh, int y_dim,
                     const double* const* factors_dev_host,
                     const double* multipliers_host, double* planes_dev,
                     void* stream);
/* Overwrites the constrained boundary vertices of component planes with the
 * plan's static Dirichlet tables, later axes winning at shared edges
 * (initial_condition.py:86-89, constrained_problem.py:286-295). */
int pml_apply_dirichlet(pml_plan* plan, double* planes_dev, long long slot,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif
