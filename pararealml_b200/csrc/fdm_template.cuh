// ---------------------------------------------------------------------------
// Fixed CUDA template of the fused FDM stage kernels (sm_100a, fp64).
//
// This file is appended to a GENERATED prelude (pararealml_b200/operators/fdm/
// codegen.py) that #defines the mesh, the equation system and the pointwise
// right-hand side emitted from the SymPy system with SymPy's C printer, and is
// compiled at run time with NVRTC (-arch=sm_100a).  It replaces, for one
// (problem, integrator) pair, the reference's
//   * ThreePointCentralDifferenceMethod._derivative / _second_derivative /
//     _add_halos_along_axis   (numerical_differentiator.py:1012-1095,1188-1242)
//   * the coordinate-system algebra of gradient/hessian/divergence/curl/
//     laplacian (numerical_differentiator.py:114-725; emitted by the generator
//     in terms of the primitives below)
//   * RK4 / ExplicitMidpoint / ForwardEuler stage arithmetic and the Dirichlet
//     overwrite after every stage (numerical_integrator.py:47-132,
//     constraint.py:43-58)
//   * the LHS.Y overwrite and the LHS.Y_LAPLACIAN Jacobi solve
//     (fdm_operator.py:127-161, numerical_differentiator.py:872-927,1097-1186)
//
// Data layout: SoA planes, plane c of a state at base + c * PML_NCELLS, cells
// in C order of the mesh axes (last mesh axis contiguous).  Unused trailing
// axes have extent 1.  Boundary tables are NaN-coded (NaN = unconstrained),
// channels-last (face cell, component), one table per (axis, side).
//
// Prelude contract (all compile-time):
//   PML_NDIM, PML_C, PML_N0, PML_N1, PML_N2, PML_COORD (0 cart, 1 polar,
//   2 cylindrical, 3 spherical), PML_NEU_MASK / PML_DIR_MASK (bit axis*2+side),
//   PML_H0..2, PML_INV2H0..2, PML_INVHH0..2 (spacing constants),
//   PML_NDT / PML_NALG / PML_NLAP and the index lists PML_DT_IDX, PML_ALG_IDX,
//   PML_LAP_IDX, PML_KIND[c] (0 dt, 1 algebraic, 2 laplacian),
//   PML_PASSTHROUGH (non-dt components of stage inputs are read from y),
//   PML_BX/BY/BZ thread block shape, PML_JAC_INV_DIAG,
//   and the generated functions pml_rhs_dt(), pml_rhs_aux().
// ---------------------------------------------------------------------------

typedef long long i64;

#define PML_NCELLS ((i64)PML_N0 * (i64)PML_N1 * (i64)PML_N2)
#define PML_NAN __longlong_as_double(0x7ff8000000000000LL)

struct PmlArgs {
  const double* u;        // stencil input of this stage (C planes)
  const double* y;        // state at the start of the step (C planes)
  const double* acc_in;   // RK4 accumulator (C planes, dt components used)
  double* u_out;          // next stage input
  double* acc_out;
  double* y_next;         // trajectory slot of this step
  double* lap_rhs;        // right-hand sides of the Y_LAPLACIAN equations
  double t_eval;          // time the right-hand side is evaluated at
  double dt;
  // boundary tables per face (axis * 2 + side), already offset by the host to
  // the time slot they are needed at (dynamic conditions)
  const double* neu[6];       // Neumann values at the evaluation time
  const double* dir[6];       // Dirichlet values of this stage's output time
  const double* dir_full[6];  // Dirichlet values of t + dt (algebraic equations)
  const double* coord[3];  // vertex coordinates along each axis
  const double* aux[4];    // 1/r[i0], sin(phi)[i2], cos(phi)[i2], 1/sin(phi)[i2]
};

template <int A> struct PmlAx;
template <> struct PmlAx<0> {
  static constexpr int N = PML_N0;
  static constexpr i64 S = (i64)PML_N1 * (i64)PML_N2;
  static constexpr double H = PML_H0, INV2H = PML_INV2H0, INVHH = PML_INVHH0;
};
template <> struct PmlAx<1> {
  static constexpr int N = PML_N1;
  static constexpr i64 S = (i64)PML_N2;
  static constexpr double H = PML_H1, INV2H = PML_INV2H1, INVHH = PML_INVHH1;
};
template <> struct PmlAx<2> {
  static constexpr int N = PML_N2;
  static constexpr i64 S = 1;
  static constexpr double H = PML_H2, INV2H = PML_INV2H2, INVHH = PML_INVHH2;
};

struct PmlCell {
  int i0, i1, i2;
  i64 idx;
};

template <int A> __device__ __forceinline__ int pml_ia(int i0, int i1, int i2) {
  return A == 0 ? i0 : (A == 1 ? i1 : i2);
}

__device__ __forceinline__ i64 pml_lin(int i0, int i1, int i2) {
  return ((i64)i0 * PML_N1 + i1) * PML_N2 + i2;
}

// index of a cell within the boundary face normal to axis A
template <int A> __device__ __forceinline__ i64 pml_face(int i0, int i1, int i2) {
  return A == 0 ? (i64)i1 * PML_N2 + i2
                : (A == 1 ? (i64)i0 * PML_N2 + i2 : (i64)i0 * PML_N1 + i1);
}

#if PML_COHERENT_LOADS
// single-CTA time loop: data written earlier in the same launch is re-read, so
// loads bypass the (non-coherent) L1 and read-only paths
#define PML_LD(p) __ldcg(p)
#define PML_LD_ONCE(p) __ldcg(p)
#define PML_ST(p, v) (*(p) = (v))
#else
#define PML_LD(p) __ldg(p)
// read-once / write-once data (step-start state, accumulator, outputs) is
// streamed so that L1/L2 keep the stencil input, which is re-read ~7 times
#if PML_STREAMING
#define PML_LD_ONCE(p) __ldcs(p)
#define PML_ST(p, v) __stcs((p), (v))
#else
#define PML_LD_ONCE(p) __ldg(p)
#define PML_ST(p, v) (*(p) = (v))
#endif
#endif

template <int A, int SIDE>
__device__ __forceinline__ double pml_neu(const PmlArgs& a, int comp, int i0,
                                          int i1, int i2) {
  constexpr int f = A * 2 + SIDE;
  if (!((PML_NEU_MASK >> f) & 1)) return PML_NAN;
  return __ldg(a.neu[f] + pml_face<A>(i0, i1, i2) * PML_C + comp);
}

// ---------------------------------------------------------------------------
// stencil sources: where the primitives read the field from.  A source returns
// the value of a component at the cell c + (D0, D1, D2); the offsets are
// compile-time constants so that every stencil load is a base pointer plus an
// immediate.
// ---------------------------------------------------------------------------
// component planes in global memory
struct PmlGlobalSrc {
  const double* const* P;
  template <int D0, int D1, int D2>
  __device__ __forceinline__ double rel(int comp, const PmlCell& c) const {
    constexpr i64 off = D0 * PmlAx<0>::S + D1 * PmlAx<1>::S + D2 * PmlAx<2>::S;
    return PML_LD(P[comp] + c.idx + off);
  }
};

// one plane (the Jacobi state of a single component)
struct PmlPlaneSrc {
  const double* p;
  template <int D0, int D1, int D2>
  __device__ __forceinline__ double rel(int, const PmlCell& c) const {
    constexpr i64 off = D0 * PmlAx<0>::S + D1 * PmlAx<1>::S + D2 * PmlAx<2>::S;
    return PML_LD(p + c.idx + off);
  }
};

template <int A, int SIDE, int D0, int D1, int D2>
__device__ __forceinline__ double pml_neu_rel(const PmlArgs& a, int comp,
                                              const PmlCell& c) {
  return pml_neu<A, SIDE>(a, comp, c.i0 + D0, c.i1 + D1, c.i2 + D2);
}

// first derivative along A at the cell c + D: zero ghost cells, boundary planes
// overwritten by the Neumann value where one exists.
// IM = interior mask: bit A set means the coordinate of c along axis A is known
// not to lie on a domain face, so (for D_A == 0) the boundary handling of that
// axis compiles away and its loads are unconditional
template <int A, int IM, int D0, int D1, int D2, class SRC>
__device__ __forceinline__ double pml_d1_rel(const PmlArgs& a, const SRC& s,
                                             int comp, const PmlCell& c) {
  typedef PmlAx<A> X;
  constexpr int e0 = A == 0, e1 = A == 1, e2 = A == 2;
  constexpr int da = A == 0 ? D0 : (A == 1 ? D1 : D2);
  if (((IM >> A) & 1) && da == 0)
    return (s.template rel<D0 + e0, D1 + e1, D2 + e2>(comp, c) -
            s.template rel<D0 - e0, D1 - e1, D2 - e2>(comp, c)) * X::INV2H;
  const int ia = pml_ia<A>(c.i0, c.i1, c.i2) + da;
  const double lo =
      ia > 0 ? s.template rel<D0 - e0, D1 - e1, D2 - e2>(comp, c) : 0.0;
  const double hi = ia < X::N - 1
                        ? s.template rel<D0 + e0, D1 + e1, D2 + e2>(comp, c)
                        : 0.0;
  double d = (hi - lo) * X::INV2H;
  if (((PML_NEU_MASK >> (A * 2)) & 1) && ia == 0) {
    const double g = pml_neu_rel<A, 0, D0, D1, D2>(a, comp, c);
    if (g == g) d = g;
  }
  if (((PML_NEU_MASK >> (A * 2 + 1)) & 1) && ia == X::N - 1) {
    const double g = pml_neu_rel<A, 1, D0, D1, D2>(a, comp, c);
    if (g == g) d = g;
  }
  return d;
}

template <int A, int IM, class SRC>
__device__ __forceinline__ double pml_d1_at(const PmlArgs& a, const SRC& s,
                                            int comp, const PmlCell& c) {
  return pml_d1_rel<A, IM, 0, 0, 0>(a, s, comp, c);
}

// neighbour pair of c along A with the second-difference ghost rule:
// ghost = inner neighbour -/+ 2 h g where a Neumann value g exists, else 0
template <int A, int IM, class SRC>
__device__ __forceinline__ void pml_nb2(const PmlArgs& a, const SRC& s, int comp,
                                        const PmlCell& c, double& lo,
                                        double& hi) {
  typedef PmlAx<A> X;
  constexpr int e0 = A == 0, e1 = A == 1, e2 = A == 2;
  if ((IM >> A) & 1) {
    lo = s.template rel<-e0, -e1, -e2>(comp, c);
    hi = s.template rel<e0, e1, e2>(comp, c);
    return;
  }
  const int ia = pml_ia<A>(c.i0, c.i1, c.i2);
  if (ia > 0) {
    lo = s.template rel<-e0, -e1, -e2>(comp, c);
  } else {
    lo = 0.0;
    if ((PML_NEU_MASK >> (A * 2)) & 1) {
      const double g = pml_neu<A, 0>(a, comp, c.i0, c.i1, c.i2);
      if (g == g) lo = s.template rel<e0, e1, e2>(comp, c) + (-2.0 * X::H) * g;
    }
  }
  if (ia < X::N - 1) {
    hi = s.template rel<e0, e1, e2>(comp, c);
  } else {
    hi = 0.0;
    if ((PML_NEU_MASK >> (A * 2 + 1)) & 1) {
      const double g = pml_neu<A, 1>(a, comp, c.i0, c.i1, c.i2);
      if (g == g) hi = s.template rel<-e0, -e1, -e2>(comp, c) + (2.0 * X::H) * g;
    }
  }
}

template <int A, int IM, class SRC>
__device__ __forceinline__ double pml_d2_at(const PmlArgs& a, const SRC& s,
                                            int comp, const PmlCell& c) {
  double lo, hi;
  pml_nb2<A, IM>(a, s, comp, c, lo, hi);
  const double v = s.template rel<0, 0, 0>(comp, c);
  return ((hi - 2.0 * v) + lo) * PmlAx<A>::INVHH;
}

// mixed second derivative: constrained d/dA, then unconstrained zero-ghost d/dB
template <int A, int B, int IM, class SRC>
__device__ __forceinline__ double pml_d2m_at(const PmlArgs& a, const SRC& s,
                                             int comp, const PmlCell& c) {
  typedef PmlAx<B> X;
  const int ib = pml_ia<B>(c.i0, c.i1, c.i2);
  constexpr int e0 = B == 0, e1 = B == 1, e2 = B == 2;
  // the two points c -/+ e_B keep the coordinate of c along A, so the interior
  // knowledge about axis A carries over to their d/dA
  constexpr bool b_in = (IM >> B) & 1;
  const double lo = (b_in || ib > 0)
                        ? pml_d1_rel<A, IM, -e0, -e1, -e2>(a, s, comp, c)
                        : 0.0;
  const double hi = (b_in || ib < X::N - 1)
                        ? pml_d1_rel<A, IM, e0, e1, e2>(a, s, comp, c)
                        : 0.0;
  return (hi - lo) * X::INV2H;
}

// Dirichlet overwrite: faces in the order axis 0 lower, axis 0 upper, axis 1
// lower, ... so later faces win on shared edges (constrained_problem.py:286-295)
template <int A>
__device__ __forceinline__ double pml_dirichlet_axis(const double* const* dir,
                                                     int comp, int i0, int i1,
                                                     int i2, double v) {
  typedef PmlAx<A> X;
  const int ia = pml_ia<A>(i0, i1, i2);
  if (((PML_DIR_MASK >> (A * 2)) & 1) && ia == 0) {
    constexpr int f = A * 2;
    const double t = __ldg(dir[f] + pml_face<A>(i0, i1, i2) * PML_C + comp);
    if (t == t) v = t;
  }
  if (((PML_DIR_MASK >> (A * 2 + 1)) & 1) && ia == X::N - 1) {
    constexpr int f = A * 2 + 1;
    const double t = __ldg(dir[f] + pml_face<A>(i0, i1, i2) * PML_C + comp);
    if (t == t) v = t;
  }
  return v;
}

__device__ __forceinline__ double pml_dirichlet(const double* const* dir,
                                                int comp, const PmlCell& c,
                                                double v) {
#if PML_DIR_MASK != 0
  if (PML_NDIM >= 1) v = pml_dirichlet_axis<0>(dir, comp, c.i0, c.i1, c.i2, v);
  if (PML_NDIM >= 2) v = pml_dirichlet_axis<1>(dir, comp, c.i0, c.i1, c.i2, v);
  if (PML_NDIM >= 3) v = pml_dirichlet_axis<2>(dir, comp, c.i0, c.i1, c.i2, v);
#endif
  return v;
}

#define PML_IM_ALL ((1 << PML_NDIM) - 1)
// every axis but the contiguous (last) one: only the two edge lanes of a mesh
// row then need boundary handling
#define PML_IM_OUTER (PML_NDIM >= 2 ? (PML_IM_ALL & ~(1 << (PML_NDIM - 1))) : 0)

// bit A set: the cell is not on a face normal to axis A
__device__ __forceinline__ int pml_interior_mask(const PmlCell& c) {
  int m = 0;
  if (PML_NDIM >= 1 && c.i0 > 0 && c.i0 < PML_N0 - 1) m |= 1;
  if (PML_NDIM >= 2 && c.i1 > 0 && c.i1 < PML_N1 - 1) m |= 2;
  if (PML_NDIM >= 3 && c.i2 > 0 && c.i2 < PML_N2 - 1) m |= 4;
  return m;
}

// x / 6.0 without the compiler's division subroutine (its call serialises the
// surrounding loads): Markstein's sequence q = x c, r = x - 6 q (exact, FMA),
// q + r c is the correctly rounded quotient throughout the normal range
__device__ __forceinline__ double pml_div6(double x) {
  const double c = 1.0 / 6.0;
  const double q = x * c;
  const double r = fma(-6.0, q, x);
  return fma(r, c, q);
}

// ---------------------------------------------------------------------------
// generated right-hand sides (prelude declares, generator defines below)
// ---------------------------------------------------------------------------
PML_GENERATED_RHS

// ---------------------------------------------------------------------------
// stage bodies
// ---------------------------------------------------------------------------
enum {
  PML_FE = 0,     // y+ = c_f(y + dt f(t, y))
  PML_MID1 = 1,   // u  = c_h(y + (dt/2) f(t, y))
  PML_MID2 = 2,   // y+ = c_f(y + dt f(t + dt/2, u))
  PML_RK4_1 = 3,  // K = dt f(t, y);      u = c_h(y + K/2); acc = K
  PML_RK4_2 = 4,  // K = dt f(t+dt/2, u); u' = c_h(y + K/2); acc += 2K
  PML_RK4_3 = 5,  // K = dt f(t+dt/2, u); u' = c_f(y + K);   acc += 2K
  PML_RK4_4 = 6   // K = dt f(t+dt, u);   y+ = c_f(y + (acc + K)/6)
};

// which variant of the generated right-hand side a warp runs: 2 = every lane is
// an interior cell (branch-free, all stencil loads unconditional so they issue
// back to back), 1 = interior along all but the contiguous axis, 0 = general.
// Must be called by all 32 lanes; inactive lanes do not constrain the choice.
__device__ __forceinline__ int pml_warp_path(bool active, const PmlCell& c) {
  const int im = active ? pml_interior_mask(c) : PML_IM_ALL;
  if (__all_sync(0xffffffffu, im == PML_IM_ALL)) return 2;
  if (PML_IM_OUTER != 0 &&
      __all_sync(0xffffffffu, (im & PML_IM_OUTER) == PML_IM_OUTER))
    return 1;
  return 0;
}

template <class SRC>
__device__ __forceinline__ void pml_eval_dt(int path, const PmlArgs& a,
                                            const SRC& src, const PmlCell& c,
                                            double t, double* K) {
  if (path == 2)
    pml_rhs_dt<PML_IM_ALL>(a, src, c, t, K);
  else if (path == 1)
    pml_rhs_dt<PML_IM_OUTER>(a, src, c, t, K);
  else
    pml_rhs_dt<0>(a, src, c, t, K);
}

template <class SRC>
__device__ __forceinline__ void pml_eval_aux(int path, const PmlArgs& a,
                                             const SRC& src, const PmlCell& c,
                                             double t, double* V) {
  if (path == 2)
    pml_rhs_aux<PML_IM_ALL>(a, src, c, t, V);
  else if (path == 1)
    pml_rhs_aux<PML_IM_OUTER>(a, src, c, t, V);
  else
    pml_rhs_aux<0>(a, src, c, t, V);
}

// algebraic (LHS.Y) and Poisson (LHS.Y_LAPLACIAN) right-hand sides use the
// step-start time and state (fdm_operator.py:127-161): they are evaluated in
// the first stage, whose stencil input is exactly that state
template <class SRC>
__device__ __forceinline__ void pml_first_stage_extras(int path,
                                                       const PmlArgs& a,
                                                       const SRC& src,
                                                       const PmlCell& c) {
#if PML_NALG + PML_NLAP > 0
  double V[PML_NALG + PML_NLAP];
  pml_eval_aux(path, a, src, c, a.t_eval, V);
#pragma unroll
  for (int j = 0; j < PML_NALG; ++j) {
    const int k = PML_ALG_IDX[j];
    a.y_next[(i64)k * PML_NCELLS + c.idx] =
        pml_dirichlet(a.dir_full, k, c, V[j]);
  }
#pragma unroll
  for (int j = 0; j < PML_NLAP; ++j)
    a.lap_rhs[(i64)j * PML_NCELLS + c.idx] = V[PML_NALG + j];
#endif
}

template <int STAGE>
__device__ __forceinline__ void pml_stage_cell(const PmlArgs& a, bool active,
                                               const PmlCell& c) {
  constexpr bool first = STAGE == PML_FE || STAGE == PML_MID1 || STAGE == PML_RK4_1;
  constexpr bool last = STAGE == PML_FE || STAGE == PML_MID2 || STAGE == PML_RK4_4;

  const int path = pml_warp_path(active, c);
  if (!active) return;

  const double* P[PML_C];
#pragma unroll
  for (int k = 0; k < PML_C; ++k) {
    const bool from_y = first || (PML_PASSTHROUGH && PML_KIND[k] != 0);
    P[k] = (from_y ? a.y : a.u) + (i64)k * PML_NCELLS;
  }
  const PmlGlobalSrc src{P};

  // the pointwise operands (step-start value, accumulator) are requested
  // before the right-hand side is evaluated, so that they travel together with
  // the stencil loads instead of in a second, dependent round trip to DRAM
  // (systems with few time-stepped components: the registers are there)
  constexpr bool early = PML_NDT <= 2;
  constexpr bool uses_acc =
      STAGE == PML_RK4_2 || STAGE == PML_RK4_3 || STAGE == PML_RK4_4;
  double y_early[PML_NDT > 0 ? PML_NDT : 1], acc_early[PML_NDT > 0 ? PML_NDT : 1];
  if (early) {
#pragma unroll
    for (int j = 0; j < PML_NDT; ++j) {
      const i64 o = (i64)PML_DT_IDX[j] * PML_NCELLS + c.idx;
      y_early[j] = first ? PML_LD(P[PML_DT_IDX[j]] + c.idx) : PML_LD_ONCE(a.y + o);
      acc_early[j] = uses_acc ? PML_LD_ONCE(a.acc_in + o) : 0.0;
    }
  }

  double K[PML_NDT > 0 ? PML_NDT : 1];
  pml_eval_dt(path, a, src, c, a.t_eval, K);

#pragma unroll
  for (int j = 0; j < PML_NDT; ++j) {
    const int k = PML_DT_IDX[j];
    const i64 o = (i64)k * PML_NCELLS + c.idx;
    const double y0 = early ? y_early[j]
                            : (first ? PML_LD(P[k] + c.idx) : PML_LD_ONCE(a.y + o));
    if (STAGE == PML_FE) {
      PML_ST(a.y_next + o, pml_dirichlet(a.dir, k, c, y0 + a.dt * K[j]));
    } else if (STAGE == PML_MID1) {
      PML_ST(a.u_out + o, pml_dirichlet(a.dir, k, c, y0 + (a.dt / 2.0) * K[j]));
    } else if (STAGE == PML_MID2) {
      PML_ST(a.y_next + o, pml_dirichlet(a.dir, k, c, y0 + a.dt * K[j]));
    } else {
      const double kk = a.dt * K[j];
      if (STAGE == PML_RK4_1) {
        PML_ST(a.acc_out + o, kk);
        PML_ST(a.u_out + o, pml_dirichlet(a.dir, k, c, y0 + kk / 2.0));
      } else if (STAGE == PML_RK4_2) {
        const double ac = early ? acc_early[j] : PML_LD_ONCE(a.acc_in + o);
        PML_ST(a.acc_out + o, ac + 2.0 * kk);
        PML_ST(a.u_out + o, pml_dirichlet(a.dir, k, c, y0 + kk / 2.0));
      } else if (STAGE == PML_RK4_3) {
        const double ac = early ? acc_early[j] : PML_LD_ONCE(a.acc_in + o);
        PML_ST(a.acc_out + o, ac + 2.0 * kk);
        PML_ST(a.u_out + o, pml_dirichlet(a.dir, k, c, y0 + kk));
      } else {
        const double ac = early ? acc_early[j] : PML_LD_ONCE(a.acc_in + o);
        PML_ST(a.y_next + o, pml_dirichlet(a.dir, k, c, y0 + pml_div6(ac + kk)));
      }
    }
  }

#if PML_NALG + PML_NLAP > 0
  // non-dt components: their time derivative is zero, so every stage input is
  // the Dirichlet-constrained step-start value (fdm_operator.py:114-120,
  // numerical_integrator.py:116-131)
  if (!last && !PML_PASSTHROUGH) {
#pragma unroll
    for (int k = 0; k < PML_C; ++k) {
      if (PML_KIND[k] == 0) continue;
      const i64 o = (i64)k * PML_NCELLS + c.idx;
      a.u_out[o] = pml_dirichlet(a.dir, k, c, PML_LD(a.y + o));
    }
  }
  if (first) pml_first_stage_extras(path, a, src, c);
#endif
}

__device__ __forceinline__ bool pml_this_cell(PmlCell& c) {
#if PML_NDIM <= 1
  c.i0 = blockIdx.x * PML_BX + threadIdx.x;
  c.i1 = 0;
  c.i2 = 0;
  if (c.i0 >= PML_N0) return false;
#elif PML_NDIM == 2
  c.i1 = blockIdx.x * PML_BX + threadIdx.x;
  c.i0 = blockIdx.y * PML_BY + threadIdx.y;
  c.i2 = 0;
  if (c.i1 >= PML_N1 || c.i0 >= PML_N0) return false;
#else
  c.i2 = blockIdx.x * PML_BX + threadIdx.x;
  c.i1 = blockIdx.y * PML_BY + threadIdx.y;
  c.i0 = blockIdx.z * PML_BZ + threadIdx.z;
  if (c.i2 >= PML_N2 || c.i1 >= PML_N1 || c.i0 >= PML_N0) return false;
#endif
  c.idx = pml_lin(c.i0, c.i1, c.i2);
  return true;
}

// stage kernels: every thread walks PML_ZREP consecutive cells along axis 0
// (fewer, longer-lived thread blocks; the plane loaded as the upper neighbour
// of one cell is the centre of the next and stays in L1)
__device__ __forceinline__ bool pml_stage_cell_coords(PmlCell& c, int rep) {
#if PML_NDIM <= 1
  (void)rep;
  return pml_this_cell(c);
#elif PML_NDIM == 2
  c.i1 = blockIdx.x * PML_BX + threadIdx.x;
  c.i0 = (blockIdx.y * PML_ZREP + rep) * PML_BY + threadIdx.y;
  c.i2 = 0;
  c.idx = pml_lin(c.i0, c.i1, 0);
  return c.i1 < PML_N1 && c.i0 < PML_N0;
#else
  c.i2 = blockIdx.x * PML_BX + threadIdx.x;
  c.i1 = blockIdx.y * PML_BY + threadIdx.y;
  c.i0 = (blockIdx.z * PML_ZREP + rep) * PML_BZ + threadIdx.z;
  c.idx = pml_lin(c.i0, c.i1, c.i2);
  return c.i2 < PML_N2 && c.i1 < PML_N1 && c.i0 < PML_N0;
#endif
}

#define PML_STAGE_KERNEL(NAME, STAGE)                                      \
  extern "C" __global__ void __launch_bounds__(PML_BX* PML_BY* PML_BZ,     \
                                               PML_MIN_BLOCKS)             \
      NAME(const __grid_constant__ PmlArgs a) {                            \
    _Pragma("unroll 1") for (int rep = 0;                                  \
                             rep < (PML_NDIM <= 1 ? 1 : PML_ZREP); ++rep) { \
      PmlCell c;                                                           \
      const bool active = pml_stage_cell_coords(c, rep);                   \
      pml_stage_cell<STAGE>(a, active, c);                                 \
    }                                                                      \
  }

PML_STAGE_KERNEL(pml_stage_fe, PML_FE)
PML_STAGE_KERNEL(pml_stage_mid1, PML_MID1)
PML_STAGE_KERNEL(pml_stage_mid2, PML_MID2)
PML_STAGE_KERNEL(pml_stage_rk4_1, PML_RK4_1)
PML_STAGE_KERNEL(pml_stage_rk4_2, PML_RK4_2)
PML_STAGE_KERNEL(pml_stage_rk4_3, PML_RK4_3)
PML_STAGE_KERNEL(pml_stage_rk4_4, PML_RK4_4)

// ---------------------------------------------------------------------------
// Fused stage pairs: temporal blocking along the slowest mesh axis.
//
// One launch performs two consecutive stages (RK4 1+2, RK4 3+4 or midpoint 1+2)
// so that the intermediate stage state never reaches HBM: RK4 costs 7 C instead
// of 17 C doubles of traffic per cell-step.
//
// A thread block owns a PML_FTX x PML_FTY tile of the in-plane axes and marches
// along axis 0 over PML_FZC planes.  Per plane ("iteration" i):
//   * the TMA unit (cp.async.bulk.tensor, one box per component plane, issued
//     by one elected lane each, completion counted on an mbarrier) streams the
//     stencil input of stage A (tile + 2 halo cells; cells outside the mesh
//     arrive as zeros) into a ring of shared-memory planes, PML_FDEPTH
//     iterations ahead; for stages 3+4 also the step-start state and the RK4
//     accumulator.  No thread ever waits on a global load, and no registers
//     are tied up by prefetches;
//   * stage A is evaluated on plane i + 1 for the tile plus ONE halo cell (one
//     thread per cell) with every stencil operand read from the input ring;
//     its result goes to a 4-slot ring of shared-memory planes;
//   * stage B is evaluated on plane i - 1 for the tile's own cells, reading
//     that ring, and writes the launch's outputs to HBM.  Stage A's increment
//     and the step-start value of a cell travel from A to B (two iterations
//     later, same thread) in registers;
//   * ONE __syncthreads() per iteration: everything an iteration reads was
//     written in an earlier iteration, so stages A and B are independent
//     instruction streams the scheduler can interleave.
// Arithmetic per cell is the same sequence of operations as in the unfused
// stage kernels above (bit-identical results).
// ---------------------------------------------------------------------------
#if PML_FUSED
struct PmlFusedArgs {
  // TMA descriptors (CUtensorMap, encoded by the host per launch) of the 4-D
  // arrays [component][axis 0][axis 1][contiguous axis]: stage A's stencil
  // input, the step-start state and the accumulator (the latter two: 3+4 only)
  alignas(64) unsigned char tm_in[128];
  alignas(64) unsigned char tm_y[128];
  alignas(64) unsigned char tm_acc[128];
  PmlArgs s;        // stage A: input planes, time, table slots; all outputs
  double t_eval_b;  // stage B evaluation time
  const double* neu_b[6];  // stage B boundary tables
  const double* dir_b[6];
};

// in-plane geometry: "x" is the contiguous mesh axis, "y" axis 1 of a 3-D mesh
#if PML_NDIM == 3
#define PML_FHY 1
#define PML_FNX PML_N2
#define PML_FNY PML_N1
#else
#define PML_FHY 0
#define PML_FNX PML_N1
#define PML_FNY 1
#endif
#define PML_MW (PML_FTX + 2)              // stage-A tile (halo 1)
#define PML_MH (PML_FTY + 2 * PML_FHY)
#define PML_IW (PML_FTX + 4)              // input tile (halo 2)
#define PML_IH (PML_FTY + 4 * PML_FHY)
#define PML_MID_PLANE (PML_MW * PML_MH)
// TMA boxes land on 128-byte boundaries: component planes are padded to 16 doubles
#define PML_PAD16(n) (((n) + 15) / 16 * 16)
#define PML_IN_PLANE PML_PAD16(PML_IW * PML_IH)
#define PML_YR_PLANE PML_PAD16(PML_IW * PML_MH)  // step-start state: rows of stage A
#define PML_OWN_PLANE PML_PAD16(PML_FTX * PML_FTY)
// components held in the rings: all of them, or only the time-stepped ones when
// the others are read from the step-start state directly (PML_PASSTHROUGH)
#define PML_NRING (PML_PASSTHROUGH ? (PML_NDT > 0 ? PML_NDT : 1) : PML_C)
#define PML_FNS_IN (PML_FDEPTH + 3)       // input ring slots
#define PML_FNS_P (PML_FDEPTH + 1)        // pointwise ring slots, barriers

__device__ __forceinline__ constexpr int pml_ring_index(int comp) {
  if (!PML_PASSTHROUGH) return comp;
  int n = 0;
  for (int k = 0; k < comp; ++k) n += PML_KIND[k] == 0;
  return n;
}

// a ring of planes in shared memory: base[d + 1] points at this thread's cell
// in the slot of plane z + d, so a stencil read is base + immediate
template <int PITCH, int PLANE>
struct PmlRingSrc {
  const double* base[3];
  const double* y;  // passthrough components are read from the state itself
  template <int D0, int D1, int D2>
  __device__ __forceinline__ double rel(int comp, const PmlCell& c) const {
    if (PML_PASSTHROUGH && PML_KIND[comp] != 0) {
      constexpr i64 off =
          D0 * PmlAx<0>::S + D1 * PmlAx<1>::S + D2 * PmlAx<2>::S;
      return PML_LD(y + (i64)comp * PML_NCELLS + c.idx + off);
    }
#if PML_NDIM == 3
    return base[D0 + 1][pml_ring_index(comp) * PLANE + D1 * PITCH + D2];
#else
    return base[D0 + 1][pml_ring_index(comp) * PLANE + D1];
#endif
  }
};

enum { PML_F_RK4_12 = 0, PML_F_RK4_34 = 1, PML_F_MID = 2 };

__device__ __forceinline__ unsigned pml_smem_addr(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
// (barriers are addressed by their 32-bit shared-memory address, computed once)
__device__ __forceinline__ void pml_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)
               : "memory");
}
__device__ __forceinline__ void pml_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void pml_mbar_wait(unsigned addr, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n"
        "  .reg .pred p;\n"
        "  mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "  selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
// one box (rows x columns of one component plane) global -> shared through the
// TMA unit; cells outside the array arrive as zeros; completion (box bytes) is
// signalled on the mbarrier
__device__ __forceinline__ void pml_tma_box(unsigned dst, const void* tmap,
                                            int x, int y, int z, int comp,
                                            unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::"
      "complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"((unsigned long long)tmap), "r"(x), "r"(y), "r"(z), "r"(comp), "r"(bar)
      : "memory");
}

// in-plane part of the variant choice (see pml_warp_path): evaluated once per
// thread block, the marching coordinate only adds a block-uniform condition
__device__ __forceinline__ int pml_inplane_path(bool active, int i1, int i2) {
  bool in_all = true, in_outer = true;
#if PML_NDIM == 3
  const bool a1 = i1 > 0 && i1 < PML_N1 - 1, a2 = i2 > 0 && i2 < PML_N2 - 1;
  in_all = a1 && a2;
  in_outer = a1;
#else
  in_all = i1 > 0 && i1 < PML_N1 - 1;
  in_outer = true;
#endif
  if (__all_sync(0xffffffffu, !active || in_all)) return 2;
  if (__all_sync(0xffffffffu, !active || in_outer)) return 1;
  return 0;
}

#if PML_FUSED == 1
template <int MODE>
__device__ __forceinline__ void pml_fused_body(const PmlFusedArgs& f,
                                               double* smem,
                                               unsigned long long* bars) {
  const PmlArgs& a = f.s;
  constexpr bool first = MODE != PML_F_RK4_34;  // stage A's input is y itself
  constexpr bool pointwise = MODE == PML_F_RK4_34;  // y and acc rings in use
  constexpr int NK = PML_NDT > 0 ? PML_NDT : 1;
  constexpr int IN_SLOT = PML_NRING * PML_IN_PLANE;
  constexpr int MID_SLOT = PML_NRING * PML_MID_PLANE;
  constexpr int YR_SLOT = NK * PML_YR_PLANE;
  constexpr int ACC_SLOT = NK * PML_OWN_PLANE;
  double* in_ring = smem;
  double* mid_ring = in_ring + PML_FNS_IN * IN_SLOT;
  double* y_ring = mid_ring + PML_PAD16(4 * MID_SLOT);  // 128-byte aligned
  double* acc_ring = y_ring + PML_FNS_P * YR_SLOT;
  const int tid = threadIdx.x;

  // ---- this thread's cell of the stage-A tile (fixed while marching)
  const int mr = PML_FHY ? tid / PML_MW : 0;
  const int mc = PML_FHY ? tid - mr * PML_MW : tid;
  const int ox = blockIdx.x * PML_FTX;  // mesh coordinates of the tile origin
#if PML_NDIM == 3
  const int oy = blockIdx.y * PML_FTY;
  const int chunk = blockIdx.z;
  const int i1 = oy - 1 + mr, i2 = ox - 1 + mc;
  const bool in_tile = tid < PML_MID_PLANE;
  const bool in_plane = in_tile && i1 >= 0 && i1 < PML_N1 && i2 >= 0 && i2 < PML_N2;
  const bool owner = in_plane && mc >= 1 && mc <= PML_FTX && mr >= 1 && mr <= PML_FTY;
  const int in_cell = (mr + 1) * PML_IW + (mc + 1);
  const int own_cell = (mr - 1) * PML_FTX + (mc - 1);
#else
  const int oy = 0;
  const int chunk = blockIdx.y;
  const int i1 = ox - 1 + mc, i2 = 0;
  const bool in_tile = tid < PML_MID_PLANE;
  const bool in_plane = in_tile && i1 >= 0 && i1 < PML_N1;
  const bool owner = in_plane && mc >= 1 && mc <= PML_FTX;
  const int in_cell = mc + 1;
  const int own_cell = mc - 1;
#endif
  const int mid_cell = mr * PML_MW + mc;
  const int yr_cell = mr * PML_IW + (mc + 1);

  // ---- planes: stage B works on [zb, ze), stage A on one more plane on each
  // side, the input ring on two more
  const int zb = chunk * PML_FZC;
  const int ze = min(zb + PML_FZC, PML_N0);
  const int a_lo = max(zb - 1, 0), a_hi = min(ze, PML_N0 - 1);
  const int in_lo = max(zb - 2, 0), in_hi = min(ze + 1, PML_N0 - 1);
  const int it0 = a_lo - 1, it1 = ze;  // iterations: A(i + 1) and B(i - 1)
  // shared-memory address of the barriers (kept in a register: re-deriving it
  // reads a special register in every iteration)
  unsigned bars_s = pml_smem_addr(bars);
  asm volatile("" : "+r"(bars_s));

  // ---- this thread's TMA box (at most one): one component plane of the
  // input tile, of the step-start state or of the accumulator.  Boxes are dealt
  // out to the warps round-robin (the copy instruction takes uniform operands,
  // so a warp issues its boxes one lane at a time)
  constexpr int N_IN_BOX = PML_NRING;
  constexpr int N_Y_BOX = pointwise ? NK : 0;
  constexpr int N_BOX = N_IN_BOX + 2 * N_Y_BOX;
  constexpr int N_WARPS = PML_F_THREADS / 32;
  static_assert(N_BOX <= PML_F_THREADS, "one TMA box per thread");
  constexpr unsigned IN_BOX_BYTES = PML_IW * PML_IH * 8;
  constexpr unsigned Y_BOX_BYTES = PML_IW * PML_MH * 8;
  constexpr unsigned ACC_BOX_BYTES = PML_FTX * PML_FTY * 8;
  int job = -1;  // 0 input, 1 step-start state, 2 accumulator
  int job_comp = 0, job_x = 0, job_y = 0;
  unsigned job_ring = 0, job_slot_bytes = 0;
  const void* job_map = nullptr;
  {
    int q = (tid >> 5) + N_WARPS * (tid & 31);
    if (q < N_IN_BOX) {
      job = 0;
      // component held at ring index q
#pragma unroll
      for (int k = 0; k < PML_C; ++k)
        if ((!PML_PASSTHROUGH || PML_KIND[k] == 0) && pml_ring_index(k) == q)
          job_comp = k;
      job_x = ox - 2;
      job_y = oy - 2 * PML_FHY;
      job_ring = pml_smem_addr(in_ring) + (unsigned)q * (PML_IN_PLANE * 8);
      job_slot_bytes = IN_SLOT * 8;
      job_map = f.tm_in;
    } else if (pointwise && (q -= N_IN_BOX) < N_Y_BOX) {
      job = 1;
      job_comp = PML_DT_IDX[q];
      job_x = ox - 2;
      job_y = oy - PML_FHY;
      job_ring = pml_smem_addr(y_ring) + (unsigned)q * (PML_YR_PLANE * 8);
      job_slot_bytes = YR_SLOT * 8;
      job_map = f.tm_y;
    } else if (pointwise && (q -= N_Y_BOX) < N_Y_BOX) {
      job = 2;
      job_comp = PML_DT_IDX[q];
      job_x = ox;
      job_y = oy;
      job_ring = pml_smem_addr(acc_ring) + (unsigned)q * (PML_OWN_PLANE * 8);
      job_slot_bytes = ACC_SLOT * 8;
      job_map = f.tm_acc;
    }
  }

  // Slot numbering: input plane p lives in slot (p - it0) mod PML_FNS_IN; the
  // step-start plane p in slot (p - it0 - 1) mod PML_FNS_P, the accumulator
  // plane p in slot (p - it0 + 1) mod PML_FNS_P and iteration j uses barrier
  // (j - it0) mod PML_FNS_P -- so that in iteration j the three pointwise
  // indices coincide (one counter).
  //
  // fetch(j, ...): everything iteration j reads for the first time -- input
  // plane j + 2 (in the prologue also j and j + 1), step-start plane j + 1
  // (stage A) and accumulator plane j - 1 (stage B); sp = pointwise slot of j,
  // si = input slot of plane j + 2
  auto fetch = [&](int j, int n_in, unsigned si, unsigned sp) {
    const unsigned bar = bars_s + sp * 8u;
    if (tid == 0) {
      unsigned tx = 0;
      for (int p = j + 3 - n_in; p <= j + 2; ++p)
        if (p >= in_lo && p <= in_hi) tx += N_IN_BOX * IN_BOX_BYTES;
      if (pointwise) {
        if (j + 1 >= a_lo && j + 1 <= a_hi) tx += N_Y_BOX * Y_BOX_BYTES;
        if (j - 1 >= zb && j - 1 < ze) tx += N_Y_BOX * ACC_BOX_BYTES;
      }
      pml_mbar_expect_tx(bar, tx);
    }
    if (job == 0) {
      for (int k = 0; k < n_in; ++k) {
        const int p = j + 3 - n_in + k;
        if (p >= in_lo && p <= in_hi)
          pml_tma_box(job_ring + (si + 1 - n_in + k) * job_slot_bytes, job_map,
                      job_x, job_y, p, job_comp, bar);
      }
    } else if (job > 0) {
      const int p = job == 1 ? j + 1 : j - 1;
      const bool valid = job == 1 ? (p >= a_lo && p <= a_hi) : (p >= zb && p < ze);
      if (valid)
        pml_tma_box(job_ring + sp * job_slot_bytes, job_map, job_x, job_y, p,
                    job_comp, bar);
    }
  };
  auto wrap = [](unsigned x, unsigned n) { return x >= n ? x - n : x; };

  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < PML_FNS_P; ++k) pml_mbar_init(bars_s + k * 8u, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // prologue: the first iteration needs three input planes at once (slots
  // 0..2), the next PML_FDEPTH - 1 iterations one more each
  fetch(it0, 3, 2, 0);
#pragma unroll 1
  for (int d = 1; d < PML_FDEPTH; ++d) fetch(it0 + d, 1, 2 + d, d);

  PmlArgs b = a;  // stage B sees its own time and table slots
  b.t_eval = f.t_eval_b;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    b.neu[q] = f.neu_b[q];
    b.dir[q] = f.dir_b[q];
  }

  // loop-invariant part of the variant choice
  const int path_a_in0 = pml_inplane_path(in_plane, i1, i2);
  const int path_b_in0 = pml_inplane_path(owner, i1, i2);

  // Per-thread loop invariants are passed through empty asm statements: the
  // compiler then has to keep them in registers instead of re-deriving them
  // from threadIdx / blockIdx in every iteration (dozens of instructions)
  unsigned inv_flags = (in_plane ? 1u : 0u) | (owner ? 2u : 0u) |
                       ((unsigned)path_a_in0 << 2) | ((unsigned)path_b_in0 << 4);
  unsigned inv_coord = (unsigned)(i1 + 1) | ((unsigned)(i2 + 1) << 16);
  int in_cell_o = in_cell, mid_cell_o = mid_cell, own_cell_o = own_cell;
  asm volatile("" : "+r"(inv_flags), "+r"(inv_coord), "+r"(in_cell_o),
               "+r"(mid_cell_o), "+r"(own_cell_o));
  const int yr_cell_o = in_cell_o - PML_FHY * PML_IW;
  const i64 idx0 = pml_lin(0, in_plane ? i1 : 0, in_plane ? i2 : 0);
  // stage B's outputs at this thread's cell, plane (iteration - 1): advanced
  // by one plane per iteration
  double* out_a =
      (MODE == PML_F_RK4_12 ? b.u_out : b.y_next) + idx0 + (i64)(it0 - 1) * PmlAx<0>::S;
  double* out_b = b.acc_out + idx0 + (i64)(it0 - 1) * PmlAx<0>::S;
  unsigned s_in = 0;   // input slot of plane i
  unsigned s_p = 0;    // pointwise slot / barrier of iteration i
  unsigned phase = 0;  // parity of the barrier's current use

  // One iteration: stage B on plane i - 1, then stage A on plane i + 1.  ka /
  // ys hold stage A's increment and the step-start value of this thread's cell
  // two iterations back (plane i - 1): stage B consumes them, stage A then
  // overwrites them for iteration i + 2.  The loop below alternates between
  // two such register sets, so nothing is ever moved.
  auto iteration = [&](int i, double (&ka)[NK], double (&ys)[NK]) {
    // the slots freed by the barrier that ended iteration i - 1 are refilled
    if (i + PML_FDEPTH <= it1)
      fetch(i + PML_FDEPTH, 1, wrap(s_in + PML_FDEPTH + 2, PML_FNS_IN),
            wrap(s_p + PML_FDEPTH, PML_FNS_P));
    pml_mbar_wait(bars_s + s_p * 8u, phase);
    // ---- stage B on plane i - 1 (tile cells), operands from the stage-A ring
    {
      const int z = i - 1;
      const bool active = (inv_flags & 2u) && z >= zb && z < ze;
      const int path = (z > 0 && z < PML_N0 - 1) ? (int)((inv_flags >> 4) & 3u) : 0;
      if (active) {
        PmlCell c;
        c.i0 = z;
        c.i1 = (int)(inv_coord & 0xffffu) - 1;
        c.i2 = (int)(inv_coord >> 16) - 1;
        c.idx = idx0 + (i64)z * PmlAx<0>::S;
        PmlRingSrc<PML_MW, PML_MID_PLANE> src;
        src.y = a.y;
#pragma unroll
        for (int d = -1; d <= 1; ++d)
          src.base[d + 1] = mid_ring + (((z + d) & 3) * MID_SLOT) + mid_cell_o;
        double K[NK];
        pml_eval_dt(path, b, src, c, b.t_eval, K);
        const double* ar = acc_ring + s_p * ACC_SLOT + own_cell_o;
#pragma unroll
        for (int j = 0; j < PML_NDT; ++j) {
          const int k = PML_DT_IDX[j];
          const i64 o = (i64)k * PML_NCELLS;
          const double y0 = ys[j];
          if (MODE == PML_F_RK4_12) {
            const double kk = b.dt * K[j];
            PML_ST(out_b + o, ka[j] + 2.0 * kk);
            PML_ST(out_a + o, pml_dirichlet(b.dir, k, c, y0 + kk / 2.0));
          } else if (MODE == PML_F_RK4_34) {
            const double kk = b.dt * K[j];
            const double acc = ar[j * PML_OWN_PLANE] + 2.0 * ka[j];
            PML_ST(out_a + o,
                   pml_dirichlet(b.dir, k, c, y0 + pml_div6(acc + kk)));
          } else {
            PML_ST(out_a + o, pml_dirichlet(b.dir, k, c, y0 + b.dt * K[j]));
          }
        }
#if PML_NALG + PML_NLAP > 0
        if (MODE == PML_F_RK4_12 && !PML_PASSTHROUGH) {
#pragma unroll
          for (int k = 0; k < PML_C; ++k) {
            if (PML_KIND[k] == 0) continue;
            const i64 o = (i64)k * PML_NCELLS + c.idx;
            b.u_out[o] = pml_dirichlet(b.dir, k, c, PML_LD(b.y + o));
          }
        }
#endif
      }
    }
    // ---- stage A on plane i + 1 (tile + halo 1), operands from the input ring
    {
      const int z = i + 1;
      const bool active = (inv_flags & 1u) && z >= a_lo && z <= a_hi;
      const int path = (z > 0 && z < PML_N0 - 1) ? (int)((inv_flags >> 2) & 3u) : 0;
      if (active) {
        PmlCell c;
        c.i0 = z;
        c.i1 = (int)(inv_coord & 0xffffu) - 1;
        c.i2 = (int)(inv_coord >> 16) - 1;
        c.idx = idx0 + (i64)z * PmlAx<0>::S;
        PmlRingSrc<PML_IW, PML_IN_PLANE> src;
        src.y = a.y;
#pragma unroll
        for (int d = 0; d < 3; ++d)
          src.base[d] =
              in_ring + wrap(s_in + d, PML_FNS_IN) * IN_SLOT + in_cell_o;
        double K[NK];
        pml_eval_dt(path, a, src, c, a.t_eval, K);
        double* slot = mid_ring + ((z & 3) * MID_SLOT) + mid_cell_o;
        const double* yr = y_ring + s_p * YR_SLOT + yr_cell_o;
#pragma unroll
        for (int j = 0; j < PML_NDT; ++j) {
          const int k = PML_DT_IDX[j];
          const double y0 = first ? src.template rel<0, 0, 0>(k, c)
                                  : yr[j * PML_YR_PLANE];
          double ua, kk = 0.0;
          if (MODE == PML_F_MID) {
            ua = y0 + (a.dt / 2.0) * K[j];
          } else {
            kk = a.dt * K[j];
            ua = MODE == PML_F_RK4_12 ? y0 + kk / 2.0 : y0 + kk;
          }
          slot[pml_ring_index(k) * PML_MID_PLANE] =
              pml_dirichlet(a.dir, k, c, ua);
          ka[j] = kk;
          ys[j] = y0;
        }
#if PML_NALG + PML_NLAP > 0
        if (!PML_PASSTHROUGH) {
#pragma unroll
          for (int k = 0; k < PML_C; ++k) {
            if (PML_KIND[k] == 0) continue;
            slot[k * PML_MID_PLANE] = pml_dirichlet(
                a.dir, k, c, PML_LD(a.y + (i64)k * PML_NCELLS + c.idx));
          }
        }
        if (first && (inv_flags & 2u) && z >= zb && z < ze)
          pml_first_stage_extras(path, a, src, c);
#endif
      }
    }
    out_a += PmlAx<0>::S;
    out_b += PmlAx<0>::S;
    s_in = wrap(s_in + 1, PML_FNS_IN);
    if (++s_p == PML_FNS_P) {
      s_p = 0;
      phase ^= 1u;
    }
    __syncthreads();
  };

  double ka_e[NK], ys_e[NK], ka_o[NK], ys_o[NK];
#pragma unroll
  for (int j = 0; j < NK; ++j) ka_e[j] = ys_e[j] = ka_o[j] = ys_o[j] = 0.0;
#pragma unroll 1
  for (int i = it0; i <= it1; i += 2) {
    iteration(i, ka_e, ys_e);
    if (i + 1 <= it1) iteration(i + 1, ka_o, ys_o);
  }
}

#define PML_FUSED_KERNEL(NAME, MODE)                                         \
  extern "C" __global__ void __launch_bounds__(PML_F_THREADS, PML_FMIN_BLOCKS) \
      NAME(const __grid_constant__ PmlFusedArgs f) {                        \
    extern __shared__ __align__(128) double pml_ring[];                      \
    __shared__ unsigned long long pml_bars[PML_FNS_P];                       \
    pml_fused_body<MODE>(f, pml_ring, pml_bars);                             \
  }

PML_FUSED_KERNEL(pml_fused_rk4_12, PML_F_RK4_12)
PML_FUSED_KERNEL(pml_fused_rk4_34, PML_F_RK4_34)
PML_FUSED_KERNEL(pml_fused_mid, PML_F_MID)
#endif  // PML_FUSED == 1
#endif  // PML_FUSED

// ---------------------------------------------------------------------------
// Fused stage pairs, warp-specialised (PML_FUSED == 2): the same temporal
// blocking as above, organised as a three-role pipeline inside one thread
// block, coupled only by mbarriers (no __syncthreads in the plane loop):
//
//   loader warp   one lane issues the TMA boxes of every plane of stage A's
//                 stencil input into a shared-memory ring, as soon as the
//                 consumers have released the slot (the pointwise operands --
//                 step-start value, accumulator -- are plain coalesced loads
//                 of the compute warps, issued ahead of their barrier waits);
//   A warps       one warp per 32 cells of a row of the stage-A tile (tile +
//                 halo 1): stage A on plane p from the input ring, result to
//                 the "mid" ring, its increment K to the "k" ring;
//   B warps       one warp per 32 cells of a row of the tile: stage B on plane
//                 p from the mid ring (needs planes p-1..p+1 of stage A),
//                 outputs straight to HBM.
//
// Both compute roles march along axis 0 and keep the three values of their own
// cell column (planes p-1, p, p+1) in registers, so a 7-point stencil costs 5
// shared-memory loads per component instead of 7.  Every role runs at its own
// pace: the A warps may lead the B warps by PML_WS_SMID - 2 planes, the loader
// the A warps by the depth of the input ring.  Plane bookkeeping is in "steps"
// r = plane - (zb - 2) of the chunk [zb, ze): the loader handles r = 0 ..
// nB + 3, stage A r = 1 .. nB + 2, stage B r = 2 .. nB + 1; planes outside the
// mesh are no-ops that still signal, so barrier slot r & 7 is in phase r >> 3
// for every role.
//   full[r & 7]   TMA bytes of loader step r (input plane r) have landed
//   adone[r & 7]  every A warp has finished plane r
//   bdone[r & 7]  every B warp has finished plane r
// ---------------------------------------------------------------------------
#if PML_FUSED == 2
#define PML_WPR (PML_MW / 32)            // warps per row of the stage-A tile
#define PML_NAW (PML_WPR * PML_MH)       // stage-A warps
#define PML_NBW (PML_WPR * PML_FTY)      // stage-B warps
#define PML_WS_THREADS (32 * (1 + PML_NAW + PML_NBW))
static_assert(PML_MW % 32 == 0, "stage-A tile rows are whole warps");
static_assert(PML_WS_THREADS == PML_F_THREADS, "thread count of the plan");

// stencil source of a marching warp: the cell's own column (planes p-1, p,
// p+1) is in registers, everything else is read from the ring
template <int PITCH, int PLANE>
struct PmlMarchSrc {
  const double* base[3];
  const double* y;  // passthrough components are read from the state itself
  double v[PML_NRING][3];
  template <int D0, int D1, int D2>
  __device__ __forceinline__ double rel(int comp, const PmlCell& c) const {
    if (PML_PASSTHROUGH && PML_KIND[comp] != 0) {
      constexpr i64 off =
          D0 * PmlAx<0>::S + D1 * PmlAx<1>::S + D2 * PmlAx<2>::S;
      return PML_LD(y + (i64)comp * PML_NCELLS + c.idx + off);
    }
    if (D1 == 0 && D2 == 0) return v[pml_ring_index(comp)][D0 + 1];
#if PML_NDIM == 3
    return base[D0 + 1][pml_ring_index(comp) * PLANE + D1 * PITCH + D2];
#else
    return base[D0 + 1][pml_ring_index(comp) * PLANE + D1];
#endif
  }
};

__device__ __forceinline__ void pml_mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// waits for the phase of barrier slot (r & 7) that step / plane r completes
__device__ __forceinline__ void pml_ws_wait(unsigned bars, int r) {
  const unsigned addr = bars + ((unsigned)r & 7u) * 8u;
  const unsigned parity = ((unsigned)r >> 3) & 1u;
  unsigned ok, spins = 0;
  do {
    asm volatile(
        "{\n"
        "  .reg .pred p;\n"
        "  mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "  selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(0x989680u)  // suspend-time hint (ns)
        : "memory");
    // a protocol error must end the launch, not hang the device
    if (!ok && ++spins > (1u << 22)) __trap();
  } while (!ok);
}

template <int MODE>
__device__ __forceinline__ void pml_ws_body(const PmlFusedArgs& f, double* smem,
                                            unsigned long long* bars) {
  const PmlArgs& a = f.s;
  constexpr bool first = MODE != PML_F_RK4_34;      // stage A's input is y itself
  constexpr bool pointwise = MODE == PML_F_RK4_34;  // needs y and acc per cell
  constexpr int NK = PML_NDT > 0 ? PML_NDT : 1;
  constexpr int S_IN = PML_WS_SIN, S_MID = PML_WS_SMID, S_K = PML_WS_SK;
  static_assert(S_IN >= 4 && S_IN <= 8 && S_MID >= 3 && S_MID <= 8, "ring depths");
  static_assert(S_K >= S_MID - 1 && S_K <= 8, "k ring covers the A/B lead");
  constexpr int IN_SLOT = PML_NRING * PML_IN_PLANE;
  constexpr int MID_SLOT = PML_NRING * PML_MID_PLANE;
  constexpr int K_SLOT = NK * PML_OWN_PLANE;
  // the TMA destination first (128-byte aligned: planes are padded to 16 doubles)
  double* in_ring = smem;
  double* mid_ring = in_ring + S_IN * IN_SLOT;
  double* k_ring = mid_ring + S_MID * MID_SLOT;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ox = blockIdx.x * PML_FTX;  // mesh coordinates of the tile origin
#if PML_NDIM == 3
  const int oy = blockIdx.y * PML_FTY;
  const int chunk = blockIdx.z;
#else
  const int oy = 0;
  const int chunk = blockIdx.y;
#endif
  const int zb = chunk * PML_FZC;
  const int ze = min(zb + PML_FZC, PML_N0);
  const int nB = ze - zb;
  const int a_lo = max(zb - 1, 0), a_hi = min(ze, PML_N0 - 1);
  const int in_lo = max(zb - 2, 0), in_hi = min(ze + 1, PML_N0 - 1);
  const unsigned full_s = pml_smem_addr(bars);
  const unsigned adone_s = full_s + 64u, bdone_s = full_s + 128u;

  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      pml_mbar_init(full_s + k * 8u, 1);
      pml_mbar_init(adone_s + k * 8u, PML_NAW);
      pml_mbar_init(bdone_s + k * 8u, PML_NBW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // =========================== loader ====================================
  if (warp == 0) {
    if (lane != 0) return;
    constexpr unsigned IN_BOX_BYTES = PML_IW * PML_IH * 8;
    const unsigned in_s = pml_smem_addr(in_ring);
    unsigned s_in = 0;
#pragma unroll 1
    for (int r = 0; r < nB + 4; ++r) {
      const int p = zb - 2 + r;
      // the slot was last read by stage A of plane r - S_IN + 1
      if (r - S_IN + 1 >= 1) pml_ws_wait(adone_s, r - S_IN + 1);
      const unsigned bar = full_s + ((unsigned)r & 7u) * 8u;
      if (p >= in_lo && p <= in_hi) {
        pml_mbar_expect_tx(bar, PML_NRING * IN_BOX_BYTES);
#pragma unroll
        for (int k = 0; k < PML_C; ++k)
          if (!PML_PASSTHROUGH || PML_KIND[k] == 0)
            pml_tma_box(in_s + (s_in * IN_SLOT + pml_ring_index(k) * PML_IN_PLANE) * 8u,
                        f.tm_in, ox - 2, oy - 2 * PML_FHY, p, k, bar);
      } else {
        pml_mbar_arrive(bar);
      }
      if (++s_in == S_IN) s_in = 0;
    }
    return;
  }

  // ============== compute roles: this lane's cell column ===================
  const bool role_a = warp <= PML_NAW;
  const int w = role_a ? warp - 1 : warp - 1 - PML_NAW;
  const int mr = role_a ? w / PML_WPR : PML_FHY + w / PML_WPR;  // row in the A tile
  const int mc = (w % PML_WPR) * 32 + lane;
#if PML_NDIM == 3
  const int i1 = oy - 1 + mr, i2 = ox - 1 + mc;
  const bool in_plane = i1 >= 0 && i1 < PML_N1 && i2 >= 0 && i2 < PML_N2;
  const bool owner = in_plane && mc >= 1 && mc <= PML_FTX && mr >= 1 && mr <= PML_FTY;
  const int in_cell = (mr + 1) * PML_IW + (mc + 1);
  const int own_cell = (mr - 1) * PML_FTX + (mc - 1);
#else
  const int i1 = ox - 1 + mc, i2 = 0;
  const bool in_plane = i1 >= 0 && i1 < PML_N1;
  const bool owner = in_plane && mc >= 1 && mc <= PML_FTX;
  const int in_cell = mc + 1;
  const int own_cell = mc - 1;
#endif
  const int mid_cell = mr * PML_MW + mc;
  const i64 idx0 = pml_lin(0, in_plane ? i1 : 0, in_plane ? i2 : 0);

  if (role_a) {
    // ============================ stage A =================================
    const int path_in0 = pml_inplane_path(in_plane, i1, i2);
    PmlMarchSrc<PML_IW, PML_IN_PLANE> src;
    src.y = a.y;
    // stage A has no plane 0: the slot's first phase is completed here so
    // that slot r & 7 is in phase r >> 3 for every plane r
    if (lane == 0) pml_mbar_arrive(adone_s);
    // input planes 0 and 1 (steps 0, 1): the column's first two values
    pml_ws_wait(full_s, 0);
    pml_ws_wait(full_s, 1);
#pragma unroll
    for (int q = 0; q < PML_NRING; ++q) {
      src.v[q][0] = 0.0;
      src.v[q][1] = in_ring[q * PML_IN_PLANE + in_cell];
      src.v[q][2] = in_ring[IN_SLOT + q * PML_IN_PLANE + in_cell];
    }
    unsigned s_lo = 0, s_c = 1, s_hi = 2;   // input slots of planes r-1, r, r+1
    unsigned s_m = 1 % S_MID, s_k = 1 % S_K;  // slots of plane r
#pragma unroll 1
    for (int r = 1; r <= nB + 2; ++r) {
      const int p = zb - 2 + r;
      const bool active = in_plane && p >= a_lo && p <= a_hi;
      // stages 3+4: the step-start value comes straight from HBM / L2,
      // requested before the waits so that it has arrived by the time the
      // right-hand side is done
      double y_start[NK];
      if (pointwise && active) {
#pragma unroll
        for (int j = 0; j < PML_NDT; ++j)
          y_start[j] = PML_LD_ONCE(a.y + (i64)PML_DT_IDX[j] * PML_NCELLS + idx0 +
                                   (i64)p * PmlAx<0>::S);
      }
      pml_ws_wait(full_s, r + 1);
#pragma unroll
      for (int q = 0; q < PML_NRING; ++q) {
        src.v[q][0] = src.v[q][1];
        src.v[q][1] = src.v[q][2];
        src.v[q][2] = in_ring[s_hi * IN_SLOT + q * PML_IN_PLANE + in_cell];
      }
      // the mid / k slots of plane r were last read by stage B of plane
      // r - S_MID + 1 (k: r - S_K, not later)
      if (r - S_MID + 1 >= 2) pml_ws_wait(bdone_s, r - S_MID + 1);
      if (active) {
        const int path = (p > 0 && p < PML_N0 - 1) ? path_in0 : 0;
        PmlCell c;
        c.i0 = p;
        c.i1 = i1;
        c.i2 = i2;
        c.idx = idx0 + (i64)p * PmlAx<0>::S;
        src.base[0] = in_ring + s_lo * IN_SLOT + in_cell;
        src.base[1] = in_ring + s_c * IN_SLOT + in_cell;
        src.base[2] = in_ring + s_hi * IN_SLOT + in_cell;
        double K[NK];
        pml_eval_dt(path, a, src, c, a.t_eval, K);
        double* slot = mid_ring + s_m * MID_SLOT + mid_cell;
        double* kslot = k_ring + s_k * K_SLOT + own_cell;
#pragma unroll
        for (int j = 0; j < PML_NDT; ++j) {
          const int k = PML_DT_IDX[j];
          const double y0 = first ? src.template rel<0, 0, 0>(k, c) : y_start[j];
          double ua, kk = 0.0;
          if (MODE == PML_F_MID) {
            ua = y0 + (a.dt / 2.0) * K[j];
          } else {
            kk = a.dt * K[j];
            ua = MODE == PML_F_RK4_12 ? y0 + kk / 2.0 : y0 + kk;
          }
          slot[pml_ring_index(k) * PML_MID_PLANE] = pml_dirichlet(a.dir, k, c, ua);
          if (MODE != PML_F_MID && owner) kslot[j * PML_OWN_PLANE] = kk;
        }
#if PML_NALG + PML_NLAP > 0
        if (!PML_PASSTHROUGH) {
#pragma unroll
          for (int k = 0; k < PML_C; ++k) {
            if (PML_KIND[k] == 0) continue;
            slot[k * PML_MID_PLANE] = pml_dirichlet(
                a.dir, k, c, PML_LD(a.y + (i64)k * PML_NCELLS + c.idx));
          }
        }
        if (first && owner && p >= zb && p < ze)
          pml_first_stage_extras(path, a, src, c);
#endif
      }
      __syncwarp();
      if (lane == 0) pml_mbar_arrive(adone_s + ((unsigned)r & 7u) * 8u);
      s_lo = s_c;
      s_c = s_hi;
      if (++s_hi == S_IN) s_hi = 0;
      if (++s_m == S_MID) s_m = 0;
      if (++s_k == S_K) s_k = 0;
    }
    return;
  }

  // ============================== stage B ===================================
  PmlArgs b = a;  // stage B sees its own time and table slots
  b.t_eval = f.t_eval_b;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    b.neu[q] = f.neu_b[q];
    b.dir[q] = f.dir_b[q];
  }
  const int path_in0 = pml_inplane_path(owner, i1, i2);
  PmlMarchSrc<PML_MW, PML_MID_PLANE> src;
  src.y = a.y;
  if (lane == 0) {  // stage B has no planes 0 and 1 (see stage A)
    pml_mbar_arrive(bdone_s);
    pml_mbar_arrive(bdone_s + 8u);
  }
  pml_ws_wait(adone_s, 1);
  pml_ws_wait(adone_s, 2);
#pragma unroll
  for (int q = 0; q < PML_NRING; ++q) {
    src.v[q][0] = 0.0;
    src.v[q][1] = mid_ring[(1 % S_MID) * MID_SLOT + q * PML_MID_PLANE + mid_cell];
    src.v[q][2] = mid_ring[(2 % S_MID) * MID_SLOT + q * PML_MID_PLANE + mid_cell];
  }
  unsigned s_lo = 1 % S_MID, s_c = 2 % S_MID, s_hi = 3 % S_MID;  // mid slots r-1, r, r+1
  unsigned s_k = 2 % S_K;                                        // k slot of plane r
  const i64 cell0 = idx0 + (i64)zb * PmlAx<0>::S;
  const double* y_at = b.y + cell0;            // step-start state at this cell
  const double* acc_at = b.acc_in + cell0;
  double* out_a = (MODE == PML_F_RK4_12 ? b.u_out : b.y_next) + cell0;
  double* out_b = b.acc_out + cell0;
#pragma unroll 1
  for (int r = 2; r <= nB + 1; ++r) {
    const int p = zb - 2 + r;
    // pointwise operands straight from HBM / L2 (the step-start value was
    // fetched for stage A a few planes ago): requested before the wait
    double y_start[NK], acc_old[NK];
    if (owner) {
#pragma unroll
      for (int j = 0; j < PML_NDT; ++j) {
        const i64 o = (i64)PML_DT_IDX[j] * PML_NCELLS;
        y_start[j] = PML_LD_ONCE(y_at + o);
        acc_old[j] = pointwise ? PML_LD_ONCE(acc_at + o) : 0.0;
      }
    }
    pml_ws_wait(adone_s, r + 1);
#pragma unroll
    for (int q = 0; q < PML_NRING; ++q) {
      src.v[q][0] = src.v[q][1];
      src.v[q][1] = src.v[q][2];
      src.v[q][2] = mid_ring[s_hi * MID_SLOT + q * PML_MID_PLANE + mid_cell];
    }
    if (owner) {
      const int path = (p > 0 && p < PML_N0 - 1) ? path_in0 : 0;
      PmlCell c;
      c.i0 = p;
      c.i1 = i1;
      c.i2 = i2;
      c.idx = idx0 + (i64)p * PmlAx<0>::S;
      src.base[0] = mid_ring + s_lo * MID_SLOT + mid_cell;
      src.base[1] = mid_ring + s_c * MID_SLOT + mid_cell;
      src.base[2] = mid_ring + s_hi * MID_SLOT + mid_cell;
      double K[NK];
      pml_eval_dt(path, b, src, c, b.t_eval, K);
      const double* kr = k_ring + s_k * K_SLOT + own_cell;
#pragma unroll
      for (int j = 0; j < PML_NDT; ++j) {
        const int k = PML_DT_IDX[j];
        const i64 o = (i64)k * PML_NCELLS;
        const double y0 = y_start[j];
        if (MODE == PML_F_RK4_12) {
          const double kk = b.dt * K[j];
          PML_ST(out_b + o, kr[j * PML_OWN_PLANE] + 2.0 * kk);
          PML_ST(out_a + o, pml_dirichlet(b.dir, k, c, y0 + kk / 2.0));
        } else if (MODE == PML_F_RK4_34) {
          const double kk = b.dt * K[j];
          const double acc = acc_old[j] + 2.0 * kr[j * PML_OWN_PLANE];
          PML_ST(out_a + o, pml_dirichlet(b.dir, k, c, y0 + pml_div6(acc + kk)));
        } else {
          PML_ST(out_a + o, pml_dirichlet(b.dir, k, c, y0 + b.dt * K[j]));
        }
      }
#if PML_NALG + PML_NLAP > 0
      if (MODE == PML_F_RK4_12 && !PML_PASSTHROUGH) {
#pragma unroll
        for (int k = 0; k < PML_C; ++k) {
          if (PML_KIND[k] == 0) continue;
          const i64 o = (i64)k * PML_NCELLS + c.idx;
          b.u_out[o] = pml_dirichlet(b.dir, k, c, PML_LD(b.y + o));
        }
      }
#endif
    }
    __syncwarp();
    if (lane == 0) pml_mbar_arrive(bdone_s + ((unsigned)r & 7u) * 8u);
    y_at += PmlAx<0>::S;
    acc_at += PmlAx<0>::S;
    out_a += PmlAx<0>::S;
    out_b += PmlAx<0>::S;
    s_lo = s_c;
    s_c = s_hi;
    if (++s_hi == S_MID) s_hi = 0;
    if (++s_k == S_K) s_k = 0;
  }
}

#define PML_WS_KERNEL(NAME, MODE)                                             \
  extern "C" __global__ void __launch_bounds__(PML_F_THREADS, PML_FMIN_BLOCKS) \
      NAME(const __grid_constant__ PmlFusedArgs f) {                         \
    extern __shared__ __align__(128) double pml_ring[];                       \
    __shared__ unsigned long long pml_bars[24];                               \
    pml_ws_body<MODE>(f, pml_ring, pml_bars);                                 \
  }

PML_WS_KERNEL(pml_fused_rk4_12, PML_F_RK4_12)
PML_WS_KERNEL(pml_fused_rk4_34, PML_F_RK4_34)
PML_WS_KERNEL(pml_fused_mid, PML_F_MID)
#endif  // PML_FUSED == 2

// ---------------------------------------------------------------------------
// Small meshes (and ODE systems): the whole time loop in ONE thread block.
// Kernel launches cost more than a stage on a few thousand cells, so all steps
// and stages run inside a single launch with __syncthreads() between stages
// (same per-cell stage arithmetic as the multi-block kernels).
// ---------------------------------------------------------------------------
#if PML_SMALL
struct PmlSmallArgs {
  PmlArgs s;        // d_t, coordinates; s.y = state before the first step;
                    // s.neu / s.dir = slot 0 of the boundary tables
  i64 neu_stride[6];  // doubles per time slot (0 = static)
  i64 dir_stride[6];
  double* traj;     // trajectory slots
  i64 stride;
  const double* t;  // start time of every step (device)
  int n_steps;
  int integrator;   // 0 forward Euler, 1 explicit midpoint, 2 RK4
  i64 slot0;
  double* u_a;
  double* u_b;
  double* acc;
  // batched solves: thread block b integrates member b (same problem, same
  // time grid); doubles between consecutive members
  i64 y_batch_stride;
  i64 traj_batch_stride;
  i64 ws_batch_stride;
};

template <int STAGE>
__device__ __forceinline__ void pml_small_stage(const PmlArgs& a) {
  for (int base = 0; base < (int)PML_NCELLS; base += PML_SMALL_THREADS) {
    const int cell = base + (int)threadIdx.x;
    PmlCell c;
    c.idx = cell;
    c.i0 = cell / (PML_N1 * PML_N2);
    c.i1 = (cell / PML_N2) % PML_N1;
    c.i2 = cell % PML_N2;
    pml_stage_cell<STAGE>(a, cell < (int)PML_NCELLS, c);
  }
  __syncthreads();
}

__device__ __forceinline__ void pml_small_set(PmlArgs& a,
                                              const PmlSmallArgs& f,
                                              const double* u, double* u_out,
                                              double t_eval, i64 neu_slot,
                                              i64 dir_slot) {
  a.u = u;
  a.u_out = u_out;
  a.t_eval = t_eval;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    a.neu[q] = f.s.neu[q] + neu_slot * f.neu_stride[q];
    a.dir[q] = f.s.dir[q] + dir_slot * f.dir_stride[q];
  }
}

extern "C" __global__ void __launch_bounds__(PML_SMALL_THREADS)
    pml_small_run(const __grid_constant__ PmlSmallArgs g) {
  // this block's member of the batch
  PmlSmallArgs f = g;
  const i64 member = blockIdx.x;
  f.s.y = g.s.y + member * g.y_batch_stride;
  f.traj = g.traj + member * g.traj_batch_stride;
  f.u_a = g.u_a + member * g.ws_batch_stride;
  f.u_b = g.u_b + member * g.ws_batch_stride;
  f.acc = g.acc + member * g.ws_batch_stride;
  PmlArgs a = f.s;
  a.acc_in = f.acc;
  a.acc_out = f.acc;
  const double dt = a.dt, half = a.dt / 2.0;
  for (int j = 0; j < f.n_steps; ++j) {
    const double t = __ldg(f.t + j);
    const double* y = j == 0 ? f.s.y : f.traj + (i64)(j - 1) * f.stride;
    a.y = y;
    a.y_next = f.traj + (i64)j * f.stride;
    const i64 s_t = f.slot0 + 3 * (i64)j, s_h = s_t + 1, s_f = s_t + 2;
#pragma unroll
    for (int q = 0; q < 6; ++q)
      a.dir_full[q] = f.s.dir[q] + s_f * f.dir_stride[q];
    if (f.integrator == 0) {
      pml_small_set(a, f, y, nullptr, t, s_t, s_f);
      pml_small_stage<PML_FE>(a);
    } else if (f.integrator == 1) {
      pml_small_set(a, f, y, f.u_a, t, s_t, s_h);
      pml_small_stage<PML_MID1>(a);
      pml_small_set(a, f, f.u_a, nullptr, t + half, s_h, s_f);
      pml_small_stage<PML_MID2>(a);
    } else {
      pml_small_set(a, f, y, f.u_a, t, s_t, s_h);
      pml_small_stage<PML_RK4_1>(a);
      pml_small_set(a, f, f.u_a, f.u_b, t + half, s_h, s_h);
      pml_small_stage<PML_RK4_2>(a);
      pml_small_set(a, f, f.u_b, f.u_a, t + half, s_h, s_f);
      pml_small_stage<PML_RK4_3>(a);
      pml_small_set(a, f, f.u_a, nullptr, t + dt, s_f, s_f);
      pml_small_stage<PML_RK4_4>(a);
    }
  }
}
#endif  // PML_SMALL

// raw right-hand side evaluation (the NumPy-in / NumPy-out differentiator entry
// points gradient/hessian/divergence/curl/laplacian are served by this kernel)
extern "C" __global__ void __launch_bounds__(PML_BX* PML_BY* PML_BZ)
    pml_eval_rhs(const __grid_constant__ PmlArgs a) {
  PmlCell c;
  if (!pml_this_cell(c)) return;
  const double* P[PML_C];
#pragma unroll
  for (int k = 0; k < PML_C; ++k) P[k] = a.u + (i64)k * PML_NCELLS;
  double K[PML_NDT > 0 ? PML_NDT : 1];
  pml_rhs_dt<0>(a, PmlGlobalSrc{P}, c, a.t_eval, K);
#pragma unroll
  for (int j = 0; j < PML_NDT; ++j) a.u_out[(i64)j * PML_NCELLS + c.idx] = K[j];
}

// static Dirichlet values written into component planes (device-side initial
// conditions; initial_condition.py:86-89)
extern "C" __global__ void __launch_bounds__(PML_BX* PML_BY* PML_BZ)
    pml_apply_dirichlet_planes(const __grid_constant__ PmlArgs a) {
  PmlCell c;
  if (!pml_this_cell(c)) return;
  if (pml_interior_mask(c) == PML_IM_ALL) return;
#pragma unroll
  for (int k = 0; k < PML_C; ++k) {
    double* q = a.u_out + (i64)k * PML_NCELLS + c.idx;
    *q = pml_dirichlet(a.dir, k, c, *q);
  }
}

// ---------------------------------------------------------------------------
// Jacobi anti-Laplacian for the LHS.Y_LAPLACIAN components
// (numerical_differentiator.py:872-927, 1097-1186).  Component j of the Jacobi
// state belongs to y component PML_LAP_IDX[j].
// ---------------------------------------------------------------------------
#if PML_NLAP > 0
struct PmlJacobiArgs {
  PmlArgs base;            // tables: neu / dir are those of t + dt
  const double* y_hat;     // NLAP planes
  const double* rhs;       // NLAP planes
  double* y_new;           // NLAP planes
  double* partials;        // one partial sum of squares per block
  int* flags;              // [0] done: set once ||y_new - y_hat|| <= tol,
                           // [1] sweeps executed, [2] block ticket
  double tol;
};

// start: channels-last (cell, NLAP) host draw -> planes, Dirichlet applied
extern "C" __global__ void __launch_bounds__(PML_BX* PML_BY* PML_BZ)
    pml_jacobi_init(const __grid_constant__ PmlArgs a,
                    const double* __restrict__ y_init, double* __restrict__ out) {
  PmlCell c;
  if (!pml_this_cell(c)) return;
#pragma unroll
  for (int j = 0; j < PML_NLAP; ++j)
    out[(i64)j * PML_NCELLS + c.idx] = pml_dirichlet(
        a.dir, PML_LAP_IDX[j], c, y_init[c.idx * PML_NLAP + j]);
}

// cell of repetition `rep`: every thread walks PML_JREP cells along axis 0, so
// that their loads are in flight together and the block reduction is amortised
__device__ __forceinline__ bool pml_jacobi_cell(PmlCell& c, int rep) {
#if PML_NDIM <= 1
  (void)rep;
  return pml_this_cell(c);
#elif PML_NDIM == 2
  c.i1 = blockIdx.x * PML_BX + threadIdx.x;
  c.i0 = (blockIdx.y * PML_JREP + rep) * PML_BY + threadIdx.y;
  c.i2 = 0;
  c.idx = pml_lin(c.i0, c.i1, 0);
  return c.i1 < PML_N1 && c.i0 < PML_N0;
#else
  c.i2 = blockIdx.x * PML_BX + threadIdx.x;
  c.i1 = blockIdx.y * PML_BY + threadIdx.y;
  c.i0 = (blockIdx.z * PML_JREP + rep) * PML_BZ + threadIdx.z;
  c.idx = pml_lin(c.i0, c.i1, c.i2);
  return c.i2 < PML_N2 && c.i1 < PML_N1 && c.i0 < PML_N0;
#endif
}

// one Jacobi update of one cell; IM as in the stencil primitives (interior
// warps run without any boundary handling)
template <int IM>
__device__ __forceinline__ double pml_jacobi_cell_update(const PmlJacobiArgs& j,
                                                         const PmlCell& c) {
  const PmlArgs& a = j.base;
  double sq = 0.0;
#pragma unroll
  for (int q = 0; q < PML_NLAP; ++q) {
    const int comp = PML_LAP_IDX[q];
    const double* p = j.y_hat + (i64)q * PML_NCELLS;
    const PmlPlaneSrc ps{p};
    double lo, hi, acc = 0.0;
#if PML_COORD == 0
    pml_nb2<0, IM>(a, ps, comp, c, lo, hi);
    acc += (lo + hi) * PML_INVHH0;
#if PML_NDIM >= 2
    pml_nb2<1, IM>(a, ps, comp, c, lo, hi);
    acc += (lo + hi) * PML_INVHH1;
#endif
#if PML_NDIM >= 3
    pml_nb2<2, IM>(a, ps, comp, c, lo, hi);
    acc += (lo + hi) * PML_INVHH2;
#endif
    acc -= PML_LD(j.rhs + (i64)q * PML_NCELLS + c.idx);
    const double v = acc * PML_JAC_INV_DIAG;
#else
    const double r = __ldg(a.coord[0] + c.i0);
    const double r2 = r * r;
    double diag;
    pml_nb2<0, IM>(a, ps, comp, c, lo, hi);
#if PML_COORD == 3
    const double s = __ldg(a.aux[1] + c.i2), co = __ldg(a.aux[2] + c.i2);
    const double r2s2 = r2 * (s * s);
    acc += (lo + hi) / (PML_H0 * PML_H0) + (hi - lo) / (PML_H0 * r);
    pml_nb2<1, IM>(a, ps, comp, c, lo, hi);
    acc += ((lo + hi) / (PML_H1 * PML_H1)) / r2s2;
    pml_nb2<2, IM>(a, ps, comp, c, lo, hi);
    acc += ((lo + hi) / (PML_H2 * PML_H2) +
            co * (hi - lo) / (2.0 * PML_H2 * s)) / r2;
    diag = 2.0 / (PML_H0 * PML_H0) + 2.0 / ((PML_H1 * PML_H1) * r2s2) +
           2.0 / ((PML_H2 * PML_H2) * r2);
#else
    acc += (lo + hi) / (PML_H0 * PML_H0) + (hi - lo) / (2.0 * PML_H0 * r);
    pml_nb2<1, IM>(a, ps, comp, c, lo, hi);
    acc += ((lo + hi) / (PML_H1 * PML_H1)) / r2;
    diag = 2.0 / (PML_H0 * PML_H0) + 2.0 / ((PML_H1 * PML_H1) * r2);
#if PML_COORD == 2
    pml_nb2<2, IM>(a, ps, comp, c, lo, hi);
    acc += (lo + hi) / (PML_H2 * PML_H2);
    diag += 2.0 / (PML_H2 * PML_H2);
#endif
#endif
    acc -= PML_LD(j.rhs + (i64)q * PML_NCELLS + c.idx);
    const double v = acc / diag;
#endif
    const double vn = pml_dirichlet(a.dir, comp, c, v);
    j.y_new[(i64)q * PML_NCELLS + c.idx] = vn;
    const double d = vn - PML_LD(p + c.idx);
    sq += d * d;
  }
  return sq;
}

extern "C" __global__ void __launch_bounds__(PML_BX* PML_BY* PML_BZ)
    pml_jacobi_sweep(const __grid_constant__ PmlJacobiArgs j) {
  if (*(const volatile int*)j.flags) return;
  double sq = 0.0;
#pragma unroll
  for (int rep = 0; rep < (PML_NDIM <= 1 ? 1 : PML_JREP); ++rep) {
    PmlCell c;
    const bool active = pml_jacobi_cell(c, rep);
    const int path = pml_warp_path(active, c);
    if (active)
      sq += path == 2 ? pml_jacobi_cell_update<PML_IM_ALL>(j, c)
                      : pml_jacobi_cell_update<0>(j, c);
  }
  // deterministic block reduction of the squared update norm
  __shared__ double red[32];
  const int tid = (threadIdx.z * PML_BY + threadIdx.y) * PML_BX + threadIdx.x;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sq += __shfl_down_sync(0xffffffffu, sq, off);
  if ((tid & 31) == 0) red[tid >> 5] = sq;
  __syncthreads();
  __shared__ int last_block;
  const int n_blocks = (int)(gridDim.x * gridDim.y * gridDim.z);
  if (tid == 0) {
    double s = 0.0;
    const int nw = (PML_BX * PML_BY * PML_BZ + 31) / 32;
    for (int w = 0; w < nw; ++w) s += red[w];
    const i64 b = ((i64)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    j.partials[b] = s;
    // the block that takes the last ticket finishes the norm (the sweep and
    // its convergence test are one launch; numerical_differentiator.py:917-925)
    __threadfence();
    last_block = atomicAdd(j.flags + 2, 1) == n_blocks - 1;
  }
  __syncthreads();
  if (!last_block) return;
  __threadfence();
  // fixed summation order (256 strided lanes, then a tree): the norm does not
  // depend on which block came last
  constexpr int NT = PML_BX * PML_BY * PML_BZ;
  __shared__ double lanes[256];
  for (int v = tid; v < 256; v += NT) {
    double s = 0.0;
    for (int i = v; i < n_blocks; i += 256) s += __ldcg(j.partials + i);
    lanes[v] = s;
  }
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    for (int v = tid; v < off; v += NT) lanes[v] += lanes[v + off];
    __syncthreads();
  }
  if (tid == 0) {
    j.flags[1] += 1;
    j.flags[2] = 0;
    if (!(sqrt(lanes[0]) > j.tol)) j.flags[0] = 1;
  }
}

// final: copies the converged planes into the trajectory slot
extern "C" __global__ void __launch_bounds__(PML_BX* PML_BY* PML_BZ)
    pml_jacobi_store(const double* __restrict__ y_hat, double* __restrict__ y_next) {
  PmlCell c;
  if (!pml_this_cell(c)) return;
#pragma unroll
  for (int j = 0; j < PML_NLAP; ++j)
    y_next[(i64)PML_LAP_IDX[j] * PML_NCELLS + c.idx] =
        y_hat[(i64)j * PML_NCELLS + c.idx];
}
#endif  // PML_NLAP > 0
