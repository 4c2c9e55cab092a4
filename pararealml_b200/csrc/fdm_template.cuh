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
//   PML_NEU_ZERO_MASK (faces whose Neumann table is 0.0 everywhere),
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
  // a static zero-flux face: the value is known at compile time
  if ((PML_NEU_ZERO_MASK >> f) & 1) return 0.0;
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

// hi - lo along A at an interior cell: the first derivative times 2 h (the
// all-interior instantiation of the generated right-hand sides multiplies the
// mesh constant into whatever the derivative is multiplied by)
template <int A, class SRC>
__device__ __forceinline__ double pml_d1raw_at(const SRC& s, int comp,
                                               const PmlCell& c) {
  constexpr int e0 = A == 0, e1 = A == 1, e2 = A == 2;
  return s.template rel<e0, e1, e2>(comp, c) - s.template rel<-e0, -e1, -e2>(comp, c);
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

// Cartesian Laplacian of one component: the sum over the axes of
// (lo + hi - 2 v) / h^2 with the ghost rule of pml_nb2, evaluated as ONE chain
// sum_a (lo_a + hi_a) / h_a^2 - 2 v sum_a 1 / h_a^2 (7 operations on a 3-D mesh
// instead of the 12 of three separate second derivatives; the rounding differs
// from theirs by a few ulp of the largest term, like the reference's own)
template <int IM, class SRC>
__device__ __forceinline__ double pml_lap_at(const PmlArgs& a, const SRC& s,
                                             int comp, const PmlCell& c) {
  const double v = s.template rel<0, 0, 0>(comp, c);
  double lo, hi;
  pml_nb2<0, IM>(a, s, comp, c, lo, hi);
  double acc = (lo + hi) * PML_INVHH0;
#if PML_NDIM >= 2
  pml_nb2<1, IM>(a, s, comp, c, lo, hi);
  acc = fma(lo + hi, PML_INVHH1, acc);
#endif
#if PML_NDIM >= 3
  pml_nb2<2, IM>(a, s, comp, c, lo, hi);
  acc = fma(lo + hi, PML_INVHH2, acc);
#endif
  return fma(v, PML_LAP_DIAG, acc);
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
  // planes [z_begin, z_end) of axis 0 this launch produces (the whole mesh, or
  // a part of it when a slab-decomposed caller launches the planes next to its
  // neighbours first); thread blocks cut the range into chunks of PML_FZC
  int z_begin, z_end;
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

// PML_F_FE2: two consecutive forward Euler steps (marching variant only): stage
// A is step j (written to u_out as well), stage B step j + 1
enum { PML_F_RK4_12 = 0, PML_F_RK4_34 = 1, PML_F_MID = 2, PML_F_FE2 = 3 };

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
__device__ __forceinline__ void pml_mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
  const int zb = f.z_begin + chunk * PML_FZC;
  const int ze = min(zb + PML_FZC, f.z_end);
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
// Fused stage pairs, column-marching variant (PML_FUSED == 2): the same
// temporal blocking and the same TMA / mbarrier plane rings as above, but the
// arithmetic is organised around registers instead of shared memory:
//
//   * a thread owns PML_FROWS consecutive rows (axis 1) of one column of the
//     stage-A tile and marches along axis 0.  The three planes z-1, z, z+1 of
//     its own cells are held in registers for BOTH stages: stage A's input
//     column is refilled with one shared-memory load per cell and plane, and
//     stage B's input column is stage A's own result, which never has to be
//     read back.  What is left for shared memory are the in-plane neighbours:
//     a 7-point stencil costs (3 R + 2) / R loads per component in stage A and
//     (2 R + 2) / R in stage B (R = rows per thread) instead of 7;
//   * the plane loop is unrolled three times, so the register columns rotate
//     by renaming, never by moves; all ring addresses, predicates and
//     boundary-variant choices are loop invariants or per-iteration scalars
//     instead of per-cell work;
//   * warps whose cells are all interior run a branch-free body for all of
//     their rows at once (independent cells interleave in the fp64 pipe);
//     only the stores are predicated.
//
// Iteration i (one __syncthreads() each): stage B on plane i - 1, stage A on
// plane i + 1 -- exactly the schedule of the variant above, and per cell the
// same sequence of arithmetic operations.
// ---------------------------------------------------------------------------
#if PML_FUSED == 2
#define PML_WPR (PML_MW / 32)                 // warps per row of the stage-A tile
#define PML_R (PML_FHY ? PML_FROWS : 1)       // rows per thread
#define PML_NRG (PML_MH / PML_R)              // row groups
static_assert(PML_MW % 32 == 0, "stage-A tile rows are whole warps");
static_assert(PML_MH % PML_R == 0, "row groups tile the stage-A tile");
static_assert(PML_F_THREADS == 32 * PML_WPR * PML_NRG, "thread count of the plan");

// stencil source of a marching thread: its own cells (PML_R rows, planes
// p-1, p, p+1; plane p + D0 sits at position (PH + D0 + 1) % 3) are
// registers, everything else is read from the ring
template <int PH, int ROW, int PITCH, int PLANE>
struct PmlColSrc {
  const double (&col)[PML_R][PML_NRING][3];
  const double* base[3];  // this thread's row-0 cell in the slots of p-1, p, p+1
  const double* y;        // passthrough components are read from the state itself
  template <int D0, int D1, int D2>
  __device__ __forceinline__ double rel(int comp, const PmlCell& c) const {
    static_assert(D0 >= -1 && D0 <= 1, "three planes");
    if (PML_PASSTHROUGH && PML_KIND[comp] != 0) {
      constexpr i64 off =
          D0 * PmlAx<0>::S + D1 * PmlAx<1>::S + D2 * PmlAx<2>::S;
      return PML_LD(y + (i64)comp * PML_NCELLS + c.idx + off);
    }
#if PML_NDIM == 3
    constexpr int row = ROW + D1;
    constexpr bool own = D2 == 0 && row >= 0 && row < PML_R;
    if constexpr (own) {
      return col[row][pml_ring_index(comp)][(PH + D0 + 1) % 3];
    } else {
      return base[D0 + 1][pml_ring_index(comp) * PLANE + row * PITCH + D2];
    }
#else
    if constexpr (D1 == 0) {
      return col[0][pml_ring_index(comp)][(PH + D0 + 1) % 3];
    } else {
      return base[D0 + 1][pml_ring_index(comp) * PLANE + D1];
    }
#endif
  }
};

template <int MODE>
struct PmlMarch {
  static constexpr bool first = MODE != PML_F_RK4_34;      // stage A's input is y itself
  static constexpr bool pointwise = MODE == PML_F_RK4_34;  // y and acc rings in use
  static constexpr int NK = PML_NDT > 0 ? PML_NDT : 1;
  static constexpr int R = PML_R;
  static constexpr int D = PML_FDEPTH;  // iterations the TMA copies run ahead
  // Synchronisation of the warps of a block.  PML_FSYNC == 0: one
  // __syncthreads() per iteration.  PML_FSYNC == 1: every warp signals the end
  // of its iteration on an mbarrier and waits, at the top of iteration n, for
  // all warps to have finished iteration n - LAG: with a 7-point stencil stage
  // B of plane i - 1 reads its neighbours' stage-A results of that plane only,
  // which were written two iterations ago (LAG = 2: the warps may drift apart
  // by a whole iteration before anyone waits); mixed derivatives also read the
  // neighbours' plane i, written in the previous iteration (LAG = 1).  Every
  // TMA-fed ring is then one slot deeper, because a slot may only be refilled
  // once all warps are two iterations past its last use.
  static constexpr int SLACK = PML_FSYNC ? 1 : 0;
  static constexpr int LAG = PML_MIXED ? 1 : 2;
  // ring depths: stage A of iteration i reads input planes i .. i + 2 (stages
  // 1+2 / midpoint: stage B re-reads the step-start value of plane i - 1 from
  // it as well), D more planes are in flight
  static constexpr int NS_IN = (first ? D + 4 : D + 3) + SLACK;
  static constexpr int NS_Y = D + 3 + SLACK;    // step-start planes i - 1 .. i + 1
  static constexpr int NS_ACC = D + 1 + SLACK;  // accumulator plane i - 1
  static constexpr int NB = D + 1 + SLACK;      // one barrier per iteration in flight
  static constexpr int IN_SLOT = PML_NRING * PML_IN_PLANE;
  static constexpr int MID_SLOT = PML_NRING * PML_MID_PLANE;
  static constexpr int YR_SLOT = NK * PML_YR_PLANE;
  static constexpr int ACC_SLOT = NK * PML_OWN_PLANE;
  static constexpr unsigned IN_BOX_BYTES = PML_IW * PML_IH * 8;
  static constexpr unsigned Y_BOX_BYTES = PML_IW * PML_MH * 8;
  static constexpr unsigned ACC_BOX_BYTES = PML_FTX * PML_FTY * 8;
  static constexpr int N_IN_BOX = PML_NRING;
  static constexpr int N_Y_BOX = pointwise ? NK : 0;

  const PmlFusedArgs& f;
  PmlArgs b;  // stage B's view: its own time and table slots
  double *in_ring, *y_ring, *acc_ring, *mid_ring;
  unsigned bars_s;
  // planes of the chunk
  int zb, ze, a_lo, a_hi, in_lo, in_hi, it0, it1;
  // this thread's column
  int i1_0, i2;
  unsigned in_mask, own_mask;  // bit r: row r is inside the mesh / a tile cell
  bool all_tile_rows;          // none of this warp's rows is a halo row
  int path_a_in, path_b_in;    // in-plane part of the variant choice
  int in_cell, mid_cell, yr_cell, own_cell;
  i64 idx_b;                   // global cell of row 0 on plane i - 1
  // this thread's TMA box (at most one per iteration)
  int job, job_comp, job_x, job_y;
  unsigned job_ring, job_slot_bytes;
  const void* job_map;
  // ring positions of iteration i (advanced once per iteration)
  unsigned s_in;   // input slot of plane i - 1
  unsigned s_y;    // step-start slot of plane i - 1
  unsigned s_acc;  // accumulator slot of plane i - 1
  unsigned s_bar, phase;
  unsigned n_it;   // iterations done (PML_FSYNC: slot / parity of the done barriers)
  // register columns (see PmlColSrc) and stage A's increments K(p) at
  // position (p - it0) % 3
  double in_col[PML_R][PML_NRING][3];
  double mid_col[PML_R][PML_NRING][3];
  double ka[PML_R][NK][3];

  __device__ __forceinline__ PmlMarch(const PmlFusedArgs& f_) : f(f_) {}

  static __device__ __forceinline__ unsigned wrap(unsigned x, unsigned n) {
    return x >= n ? x - n : x;
  }

  // everything iteration j reads for the first time: input plane j + 2 (the
  // first iteration also j and j + 1), step-start plane j + 1 (stage A) and
  // accumulator plane j - 1 (stage B)
  __device__ __forceinline__ void fetch(int j, int n_in, unsigned si, unsigned sy,
                                        unsigned sa, unsigned sb) {
    const unsigned bar = bars_s + sb * 8u;
    if (threadIdx.x == 0) {
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
      // si: slot of plane j + 2
      for (int k = 0; k < n_in; ++k) {
        const int p = j + 3 - n_in + k;
        if (p >= in_lo && p <= in_hi)
          pml_tma_box(job_ring + wrap(si + NS_IN + 1 - n_in + k, NS_IN) * job_slot_bytes,
                      job_map, job_x, job_y, p, job_comp, bar);
      }
    } else if (job == 1) {
      if (j + 1 >= a_lo && j + 1 <= a_hi)
        pml_tma_box(job_ring + sy * job_slot_bytes, job_map, job_x, job_y, j + 1,
                    job_comp, bar);
    } else if (job == 2) {
      if (j - 1 >= zb && j - 1 < ze)
        pml_tma_box(job_ring + sa * job_slot_bytes, job_map, job_x, job_y, j - 1,
                    job_comp, bar);
    }
  }

  __device__ __forceinline__ void setup(double* smem, unsigned long long* bars) {
    const PmlArgs& a = f.s;
    in_ring = smem;  // TMA destinations first: 128-byte aligned planes
    y_ring = in_ring + NS_IN * IN_SLOT;
    acc_ring = y_ring + (pointwise ? NS_Y * YR_SLOT : 0);
    mid_ring = acc_ring + (pointwise ? NS_ACC * ACC_SLOT : 0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int strip = PML_FHY ? warp % PML_WPR : warp;
    const int mr0 = PML_FHY ? (warp / PML_WPR) * R : 0;
    const int mc = strip * 32 + lane;
    const int ox = blockIdx.x * PML_FTX;  // mesh coordinates of the tile origin
#if PML_NDIM == 3
    const int oy = blockIdx.y * PML_FTY;
    const int chunk = blockIdx.z;
    i1_0 = oy - 1 + mr0;
    i2 = ox - 1 + mc;
    in_mask = own_mask = 0;
    bool all_in = true, outer_in = true, all_in_b = true, outer_in_b = true;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i1 = i1_0 + r;
      const bool inp = i1 >= 0 && i1 < PML_N1 && i2 >= 0 && i2 < PML_N2;
      const bool own = inp && mc >= 1 && mc <= PML_FTX && mr0 + r >= 1 &&
                       mr0 + r <= PML_FTY;
      in_mask |= (inp ? 1u : 0u) << r;
      own_mask |= (own ? 1u : 0u) << r;
      const bool a1 = i1 > 0 && i1 < PML_N1 - 1, a2 = i2 > 0 && i2 < PML_N2 - 1;
      all_in = all_in && (!inp || (a1 && a2));
      outer_in = outer_in && (!inp || a1);
      all_in_b = all_in_b && (!own || (a1 && a2));
      outer_in_b = outer_in_b && (!own || a1);
    }
    in_cell = (mr0 + 1) * PML_IW + (mc + 1);
    own_cell = (mr0 - 1) * PML_FTX + (mc - 1);
    idx_b = (i64)i1_0 * PML_N2 + i2;
    all_tile_rows = mr0 >= 1 && mr0 + R - 1 <= PML_FTY;
#else
    const int oy = 0;
    const int chunk = blockIdx.y;
    i1_0 = ox - 1 + mc;
    i2 = 0;
    const bool inp = i1_0 >= 0 && i1_0 < PML_N1;
    const bool own = inp && mc >= 1 && mc <= PML_FTX;
    in_mask = inp ? 1u : 0u;
    own_mask = own ? 1u : 0u;
    const bool a1 = i1_0 > 0 && i1_0 < PML_N1 - 1;
    const bool all_in = !inp || a1, outer_in = true;
    const bool all_in_b = !own || a1, outer_in_b = true;
    in_cell = mc + 1;
    own_cell = mc - 1;
    idx_b = (i64)i1_0;
    all_tile_rows = true;
#endif
    mid_cell = mr0 * PML_MW + mc;
    yr_cell = mr0 * PML_IW + (mc + 1);
    path_a_in = __all_sync(0xffffffffu, all_in) ? 2
                : (__all_sync(0xffffffffu, outer_in) ? 1 : 0);
    path_b_in = __all_sync(0xffffffffu, all_in_b) ? 2
                : (__all_sync(0xffffffffu, outer_in_b) ? 1 : 0);
    // components that are not time-stepped are read from the state at the
    // cell itself: rows outside the mesh must not be evaluated then, which
    // only the guarded variants ensure
    if (PML_NALG + PML_NLAP > 0) {
      if (__any_sync(0xffffffffu, in_mask != (1u << R) - 1u)) path_a_in = min(path_a_in, 1);
      if (__any_sync(0xffffffffu, own_mask != (1u << R) - 1u)) path_b_in = min(path_b_in, 1);
    }

    zb = f.z_begin + chunk * PML_FZC;
    ze = min(zb + PML_FZC, f.z_end);
    a_lo = max(zb - 1, 0);
    a_hi = min(ze, PML_N0 - 1);
    in_lo = max(zb - 2, 0);
    in_hi = min(ze + 1, PML_N0 - 1);
    it0 = a_lo - 1;
    it1 = ze;  // iterations: A(i + 1) and B(i - 1)
    idx_b += (i64)(it0 - 1) * PmlAx<0>::S;

    b = a;
    b.t_eval = f.t_eval_b;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      b.neu[q] = f.neu_b[q];
      b.dir[q] = f.dir_b[q];
    }

    // TMA boxes are dealt out to the warps round-robin (the copy instruction
    // takes uniform operands, so a warp issues its boxes one lane at a time)
    constexpr int N_WARPS = PML_F_THREADS / 32;
    static_assert(N_IN_BOX + 2 * N_Y_BOX <= PML_F_THREADS, "one TMA box per thread");
    job = -1;
    job_comp = job_x = job_y = 0;
    job_ring = job_slot_bytes = 0;
    job_map = nullptr;
    int q = warp + N_WARPS * lane;
    if (q < N_IN_BOX) {
      job = 0;
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

    bars_s = pml_smem_addr(bars);
    if (tid == 0) {
#pragma unroll
      for (int k = 0; k < NB; ++k) pml_mbar_init(bars_s + k * 8u, 1);
      // "iteration done" barriers (PML_FSYNC): one arrival per warp
#pragma unroll
      for (int k = 0; k < 4; ++k) pml_mbar_init(bars_s + (NB + k) * 8u, N_WARPS);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // Slot numbering: input plane p lives in slot (p - it0 + 1) mod NS_IN, the
    // step-start plane p in slot (p - it0 + 1) mod NS_Y, the accumulator plane
    // p in slot (p - it0 + 1) mod NS_ACC (so that "plane i - 1" is the
    // iteration's running index in all three) and iteration j uses barrier
    // (j - it0) mod NB.
    // prologue: the first iteration needs three input planes at once, the
    // next D - 1 iterations one more each
    fetch(it0, 3, 3 % NS_IN, 2 % NS_Y, 0, 0);
#pragma unroll 1
    for (int d = 1; d < D; ++d)
      fetch(it0 + d, 1, (3 + d) % NS_IN, (2 + d) % NS_Y, d % NS_ACC, d % NB);
    s_in = s_y = s_acc = s_bar = phase = 0;
    n_it = 0;

    // the column's first two planes (it0, it0 + 1: positions 0, 1 of phase 0)
    pml_mbar_wait(bars_s, 0);
    const double* p0 = in_ring + (1 % NS_IN) * IN_SLOT + in_cell;
    const double* p1 = in_ring + (2 % NS_IN) * IN_SLOT + in_cell;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int q2 = 0; q2 < PML_NRING; ++q2) {
        in_col[r][q2][0] = p0[q2 * PML_IN_PLANE + r * PML_IW];
        in_col[r][q2][1] = p1[q2 * PML_IN_PLANE + r * PML_IW];
        in_col[r][q2][2] = 0.0;
        mid_col[r][q2][0] = mid_col[r][q2][1] = mid_col[r][q2][2] = 0.0;
      }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int j = 0; j < NK; ++j) ka[r][j][0] = ka[r][j][1] = ka[r][j][2] = 0.0;
  }

  // ---- stage A on plane p = i + 1, row ROW of this thread
  template <int PH, int ROW, int IM>
  __device__ __forceinline__ void a_row(int p, i64 idx_row0, const double* pin_lo,
                                        const double* pin_c, const double* pin_hi,
                                        const double* py, double* pm) {
    const PmlArgs& a = f.s;
    PmlCell c;
    c.i0 = p;
    c.i1 = PML_FHY ? i1_0 + ROW : i1_0;
    c.i2 = i2;
    c.idx = idx_row0 + (PML_FHY ? (i64)ROW * PmlAx<1>::S : 0);
    const PmlColSrc<PH, ROW, PML_IW, PML_IN_PLANE> src{in_col, {pin_lo, pin_c, pin_hi}, a.y};
    double K[NK];
    pml_rhs_dt<IM>(a, src, c, a.t_eval, K);
#pragma unroll
    for (int j = 0; j < PML_NDT; ++j) {
      const int k = PML_DT_IDX[j];
      const double y0 = first ? in_col[ROW][pml_ring_index(k)][(PH + 1) % 3]
                              : py[j * PML_YR_PLANE + ROW * PML_IW];
      double ua, kk = 0.0;
      if (MODE == PML_F_MID) {
        ua = y0 + (a.dt / 2.0) * K[j];
      } else if (MODE == PML_F_FE2) {
        ua = y0 + a.dt * K[j];
      } else {
        kk = a.dt * K[j];
        ua = MODE == PML_F_RK4_12 ? y0 + kk / 2.0 : y0 + kk;
      }
      const double v = pml_dirichlet(a.dir, k, c, ua);
      mid_col[ROW][pml_ring_index(k)][PH % 3] = v;
      pm[pml_ring_index(k) * PML_MID_PLANE + ROW * PML_MW] = v;
      if (MODE == PML_F_FE2) {
        // the first of the two steps is a result as well
        if (((own_mask >> ROW) & 1u) && p >= zb && p < ze)
          PML_ST(a.u_out + (i64)k * PML_NCELLS + c.idx, v);
      } else {
        ka[ROW][j][(PH + 1) % 3] = kk;
      }
    }
#if PML_NALG + PML_NLAP > 0
    if (!PML_PASSTHROUGH) {
#pragma unroll
      for (int k = 0; k < PML_C; ++k) {
        if (PML_KIND[k] == 0) continue;
        const double v = pml_dirichlet(
            a.dir, k, c, PML_LD(a.y + (i64)k * PML_NCELLS + c.idx));
        mid_col[ROW][k][PH % 3] = v;
        pm[k * PML_MID_PLANE + ROW * PML_MW] = v;
      }
    }
    if (first && ((own_mask >> ROW) & 1u) && p >= zb && p < ze)
      pml_first_stage_extras(IM == PML_IM_ALL ? 2 : (IM == 0 ? 0 : 1), a, src, c);
#endif
  }

  template <int PH, int IM, bool GUARD, int ROW = 0>
  __device__ __forceinline__ void a_rows(int p, i64 idx_row0, const double* pin_lo,
                                         const double* pin_c, const double* pin_hi,
                                         const double* py, double* pm) {
    if constexpr (ROW < R) {
      if (!GUARD || ((in_mask >> ROW) & 1u))
        a_row<PH, ROW, IM>(p, idx_row0, pin_lo, pin_c, pin_hi, py, pm);
      a_rows<PH, IM, GUARD, ROW + 1>(p, idx_row0, pin_lo, pin_c, pin_hi, py, pm);
    }
  }

  // ---- stage B on plane p = i - 1, row ROW of this thread
  template <int PH, int ROW, int IM>
  __device__ __forceinline__ void b_row(int p, bool store, const double* pm_lo,
                                        const double* pm_c, const double* pm_hi,
                                        const double* py, const double* pacc,
                                        double* out_a, double* out_b) {
    PmlCell c;
    c.i0 = p;
    c.i1 = PML_FHY ? i1_0 + ROW : i1_0;
    c.i2 = i2;
    c.idx = idx_b + (PML_FHY ? (i64)ROW * PmlAx<1>::S : 0);
    const PmlColSrc<PH, ROW, PML_MW, PML_MID_PLANE> src{mid_col, {pm_lo, pm_c, pm_hi}, f.s.y};
    double K[NK];
    pml_rhs_dt<IM>(b, src, c, b.t_eval, K);
#pragma unroll
    for (int j = 0; j < PML_NDT; ++j) {
      const int k = PML_DT_IDX[j];
      const i64 o = (i64)k * PML_NCELLS + c.idx;
      // the step-start value: stages 1+2 / midpoint re-read it from the input
      // ring (plane i - 1 is still there), stages 3+4 from the y ring
      // (two Euler steps: the second starts from stage A's own result)
      const double y0 = MODE == PML_F_FE2
                            ? mid_col[ROW][pml_ring_index(k)][(PH + 1) % 3]
                            : (first ? py[pml_ring_index(k) * PML_IN_PLANE + ROW * PML_IW]
                                     : py[j * PML_YR_PLANE + ROW * PML_IW]);
      const double k_a = ka[ROW][j][(PH + 2) % 3];
      if (MODE == PML_F_RK4_12) {
        const double kk = b.dt * K[j];
        const double acc = k_a + 2.0 * kk;
        const double u = pml_dirichlet(b.dir, k, c, y0 + kk / 2.0);
        if (store) {
          PML_ST(out_b + o, acc);
          PML_ST(out_a + o, u);
        }
      } else if (MODE == PML_F_RK4_34) {
        const double kk = b.dt * K[j];
        const double acc = pacc[j * PML_OWN_PLANE + ROW * PML_FTX] + 2.0 * k_a;
        const double u = pml_dirichlet(b.dir, k, c, y0 + pml_div6(acc + kk));
        if (store) PML_ST(out_a + o, u);
      } else {
        const double u = pml_dirichlet(b.dir, k, c, y0 + b.dt * K[j]);
        if (store) PML_ST(out_a + o, u);
      }
    }
#if PML_NALG + PML_NLAP > 0
    if (MODE == PML_F_RK4_12 && !PML_PASSTHROUGH && store) {
#pragma unroll
      for (int k = 0; k < PML_C; ++k) {
        if (PML_KIND[k] == 0) continue;
        const i64 o = (i64)k * PML_NCELLS + c.idx;
        b.u_out[o] = pml_dirichlet(b.dir, k, c, PML_LD(b.y + o));
      }
    }
#endif
  }

  template <int PH, int IM, bool GUARD, int ROW = 0>
  __device__ __forceinline__ void b_rows(int p, const double* pm_lo, const double* pm_c,
                                         const double* pm_hi, const double* py,
                                         const double* pacc, double* out_a,
                                         double* out_b) {
    if constexpr (ROW < R) {
      const bool own = (own_mask >> ROW) & 1u;
      if (!GUARD || own)
        b_row<PH, ROW, IM>(p, own, pm_lo, pm_c, pm_hi, py, pacc, out_a, out_b);
      b_rows<PH, IM, GUARD, ROW + 1>(p, pm_lo, pm_c, pm_hi, py, pacc, out_a, out_b);
    }
  }

  // One iteration at phase PH = (i - it0) mod 3: stage B on plane i - 1, then
  // stage A on plane i + 1.
  template <int PH>
  __device__ __forceinline__ void iteration(int i) {
    const PmlArgs& a = f.s;
    if (PML_FSYNC && n_it >= (unsigned)LAG) {
      const unsigned m = n_it - LAG;
      pml_mbar_wait(bars_s + (NB + (m & 3u)) * 8u, (m >> 2) & 1u);
    }
    // the slots whose last readers are known to be done are refilled
    if (i + D <= it1)
      fetch(i + D, 1, wrap(s_in + D + 3, NS_IN), wrap(s_y + D + 2, NS_Y),
            wrap(s_acc + D, NS_ACC), wrap(s_bar + D, NB));
    // ring addresses of this iteration
    const double* pin_b = in_ring + s_in * IN_SLOT + in_cell;  // plane i - 1
    const double* pin_lo = in_ring + wrap(s_in + 1, NS_IN) * IN_SLOT + in_cell;
    const double* pin_c = in_ring + wrap(s_in + 2, NS_IN) * IN_SLOT + in_cell;
    const double* pin_hi = in_ring + wrap(s_in + 3, NS_IN) * IN_SLOT + in_cell;
    const double* py_b = y_ring + s_y * YR_SLOT + yr_cell;                 // plane i - 1
    const double* py_a = y_ring + wrap(s_y + 2, NS_Y) * YR_SLOT + yr_cell;  // plane i + 1
    const double* pacc = acc_ring + s_acc * ACC_SLOT + own_cell;
    const double* pm_lo = mid_ring + ((i - 2) & 3) * MID_SLOT + mid_cell;
    const double* pm_c = mid_ring + ((i - 1) & 3) * MID_SLOT + mid_cell;
    const double* pm_hi = mid_ring + (i & 3) * MID_SLOT + mid_cell;
    double* pm_new = mid_ring + ((i + 1) & 3) * MID_SLOT + mid_cell;

    pml_mbar_wait(bars_s + s_bar * 8u, phase);
    // the column's new plane i + 2
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int q = 0; q < PML_NRING; ++q)
        in_col[r][q][(PH + 2) % 3] = pin_hi[q * PML_IN_PLANE + r * PML_IW];

    // ---- stage B on plane i - 1 (tile cells), neighbours from the stage-A ring
    {
      const int p = i - 1;
      if (p >= zb && p < ze) {
        const int path = (p > 0 && p < PML_N0 - 1) ? path_b_in : 0;
        double* out_a = MODE == PML_F_RK4_12 ? b.u_out : b.y_next;
        const double* py = first ? pin_b : py_b;
        if (path == 2 && all_tile_rows)
          b_rows<PH, PML_IM_ALL, false>(p, pm_lo, pm_c, pm_hi, py, pacc, out_a, b.acc_out);
        else if (path == 2)  // first / last row group: its halo row is skipped
          b_rows<PH, PML_IM_ALL, true>(p, pm_lo, pm_c, pm_hi, py, pacc, out_a, b.acc_out);
        else if (PML_F_PATH1 && path == 1)
          b_rows<PH, PML_IM_OUTER, true>(p, pm_lo, pm_c, pm_hi, py, pacc, out_a, b.acc_out);
        else
          b_rows<PH, 0, true>(p, pm_lo, pm_c, pm_hi, py, pacc, out_a, b.acc_out);
      }
    }
    // ---- stage A on plane i + 1 (tile + halo 1), neighbours from the input ring
    {
      const int p = i + 1;
      if (p >= a_lo && p <= a_hi) {
        const int path = (p > 0 && p < PML_N0 - 1) ? path_a_in : 0;
        const i64 idx_row0 = idx_b + 2 * PmlAx<0>::S;
        if (path == 2)
          a_rows<PH, PML_IM_ALL, false>(p, idx_row0, pin_lo, pin_c, pin_hi, py_a, pm_new);
        else if (PML_F_PATH1 && path == 1)
          a_rows<PH, PML_IM_OUTER, true>(p, idx_row0, pin_lo, pin_c, pin_hi, py_a, pm_new);
        else
          a_rows<PH, 0, true>(p, idx_row0, pin_lo, pin_c, pin_hi, py_a, pm_new);
      }
    }
    idx_b += PmlAx<0>::S;
    s_in = wrap(s_in + 1, NS_IN);
    s_y = wrap(s_y + 1, NS_Y);
    s_acc = wrap(s_acc + 1, NS_ACC);
    if (++s_bar == NB) {
      s_bar = 0;
      phase ^= 1u;
    }
#if PML_FSYNC
    __syncwarp();
    if ((threadIdx.x & 31) == 0)
      pml_mbar_arrive(bars_s + (NB + (n_it & 3u)) * 8u);
    ++n_it;
#else
    __syncthreads();
#endif
  }

  __device__ __forceinline__ void run() {
#pragma unroll 1
    for (int i = it0; i <= it1; i += 3) {
      iteration<0>(i);
      if (i + 1 <= it1) iteration<1>(i + 1);
      if (i + 2 <= it1) iteration<2>(i + 2);
    }
  }
};

#define PML_MARCH_KERNEL(NAME, MODE)                                          \
  extern "C" __global__ void __launch_bounds__(PML_F_THREADS, PML_FMIN_BLOCKS) \
      NAME(const __grid_constant__ PmlFusedArgs f) {                         \
    extern __shared__ __align__(128) double pml_ring[];                       \
    __shared__ unsigned long long pml_bars[PML_FDEPTH + 2 + 4];               \
    PmlMarch<MODE> m(f);                                                      \
    m.setup(pml_ring, pml_bars);                                              \
    m.run();                                                                  \
  }

PML_MARCH_KERNEL(pml_fused_rk4_12, PML_F_RK4_12)
PML_MARCH_KERNEL(pml_fused_rk4_34, PML_F_RK4_34)
PML_MARCH_KERNEL(pml_fused_mid, PML_F_MID)
PML_MARCH_KERNEL(pml_fused_fe2, PML_F_FE2)
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
__device__ __forceinline__ bool pml_jacobi_cell_of(PmlCell& c, int rep, int bx,
                                                   int by, int bz) {
#if PML_NDIM <= 1
  (void)rep; (void)by; (void)bz;
  c.i0 = bx * PML_BX + threadIdx.x;
  c.i1 = 0;
  c.i2 = 0;
  c.idx = c.i0;
  return c.i0 < PML_N0;
#elif PML_NDIM == 2
  (void)bz;
  c.i1 = bx * PML_BX + threadIdx.x;
  c.i0 = (by * PML_JREP + rep) * PML_BY + threadIdx.y;
  c.i2 = 0;
  c.idx = pml_lin(c.i0, c.i1, 0);
  return c.i1 < PML_N1 && c.i0 < PML_N0;
#else
  c.i2 = bx * PML_BX + threadIdx.x;
  c.i1 = by * PML_BY + threadIdx.y;
  c.i0 = (bz * PML_JREP + rep) * PML_BZ + threadIdx.z;
  c.idx = pml_lin(c.i0, c.i1, c.i2);
  return c.i2 < PML_N2 && c.i1 < PML_N1 && c.i0 < PML_N0;
#endif
}
__device__ __forceinline__ bool pml_jacobi_cell(PmlCell& c, int rep) {
  return pml_jacobi_cell_of(c, rep, blockIdx.x, blockIdx.y, blockIdx.z);
}

// one plane read with ordinary (coherent) loads: the persistent Jacobi loop
// re-reads planes other thread blocks wrote earlier in the same launch, which
// the read-only path of PmlPlaneSrc must not be used for
struct PmlPlaneSrcCoherent {
  const double* p;
  template <int D0, int D1, int D2>
  __device__ __forceinline__ double rel(int, const PmlCell& c) const {
    constexpr i64 off = D0 * PmlAx<0>::S + D1 * PmlAx<1>::S + D2 * PmlAx<2>::S;
    return p[c.idx + off];
  }
};

// one Jacobi update of one cell; IM as in the stencil primitives (interior
// warps run without any boundary handling)
template <int IM, class PLANE = PmlPlaneSrc>
__device__ __forceinline__ double pml_jacobi_cell_update(const PmlJacobiArgs& j,
                                                         const PmlCell& c) {
  const PmlArgs& a = j.base;
  double sq = 0.0;
#pragma unroll
  for (int q = 0; q < PML_NLAP; ++q) {
    const int comp = PML_LAP_IDX[q];
    const double* p = j.y_hat + (i64)q * PML_NCELLS;
    const PLANE ps{p};
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
    const double d = vn - ps.template rel<0, 0, 0>(comp, c);
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

// ---------------------------------------------------------------------------
// The whole Jacobi iteration in ONE cooperative launch: a persistent grid (as
// many thread blocks as are resident at once) walks the tiles of the sweep
// kernel above, sweep after sweep, with a grid-wide barrier in between; every
// block then adds up the per-block partial sums of the update norm in the same
// fixed order, so all blocks take the same decision to stop
// (numerical_differentiator.py:917-925) without another barrier or a round
// trip to the host.  flags[3] is the barrier counter (zeroed by the host).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pml_grid_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
    unsigned seen;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];"
                   : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
  }
  __syncthreads();
}

extern "C" __global__ void __launch_bounds__(PML_BX* PML_BY* PML_BZ)
    pml_jacobi_loop(const __grid_constant__ PmlJacobiArgs j0, int max_sweeps,
                    int gx, int gy, int gz) {
  PmlJacobiArgs j = j0;
  constexpr int NT = PML_BX * PML_BY * PML_BZ;
  const int tid = (threadIdx.z * PML_BY + threadIdx.y) * PML_BX + threadIdx.x;
  const int n_tiles = gx * gy * gz;
  const int n_blocks = (int)gridDim.x;
  unsigned* barrier = (unsigned*)(j.flags + 3);
  __shared__ double red[32];
  __shared__ double lanes[256];
  int sweeps = 0, done = 0;
  while (!done && (max_sweeps <= 0 || sweeps < max_sweeps)) {
    double sq = 0.0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < n_tiles; tile += n_blocks) {
      const int bx = tile % gx, by = (tile / gx) % gy, bz = tile / (gx * gy);
#pragma unroll
      for (int rep = 0; rep < (PML_NDIM <= 1 ? 1 : PML_JREP); ++rep) {
        PmlCell c;
        const bool active = pml_jacobi_cell_of(c, rep, bx, by, bz);
        const int path = pml_warp_path(active, c);
        if (active)
          sq += path == 2
                    ? pml_jacobi_cell_update<PML_IM_ALL, PmlPlaneSrcCoherent>(j, c)
                    : pml_jacobi_cell_update<0, PmlPlaneSrcCoherent>(j, c);
      }
    }
    // deterministic block reduction, one partial per block and sweep parity
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sq += __shfl_down_sync(0xffffffffu, sq, off);
    if ((tid & 31) == 0) red[tid >> 5] = sq;
    __syncthreads();
    double* partials = j.partials + (sweeps & 1) * n_blocks;
    if (tid == 0) {
      double s = 0.0;
      for (int w = 0; w < (NT + 31) / 32; ++w) s += red[w];
      partials[blockIdx.x] = s;
    }
    ++sweeps;
    pml_grid_barrier(barrier, (unsigned)sweeps * (unsigned)n_blocks);
    // the norm of this sweep's update, summed identically by every block
    for (int v = tid; v < 256; v += NT) {
      double s = 0.0;
      for (int i = v; i < n_blocks; i += 256) s += __ldcg(partials + i);
      lanes[v] = s;
    }
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
      for (int v = tid; v < off; v += NT) lanes[v] += lanes[v + off];
      __syncthreads();
    }
    done = !(sqrt(lanes[0]) > j.tol);
    __syncthreads();  // lanes / red are rewritten by the next sweep
    const double* swap = j.y_hat;
    j.y_hat = j.y_new;
    j.y_new = const_cast<double*>(swap);
  }
  if (blockIdx.x == 0 && tid == 0) {
    j.flags[0] = done;
    j.flags[1] = sweeps;
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
