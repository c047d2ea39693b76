// oracle/oracle.cpp
//
// TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's element-level
// AD assembly algorithm (dohyun-cse/mfem-ad).  Nothing in the product path
// (mfem-ad_b200/, include/) may link, import or execute this file; only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs use it, and only as the checker / CPU baseline.
//
// PARITY PINNING: the reference cannot be compiled in this image (it needs
// MFEM, MPI, hypre, UMFPACK, MUMPS -- none installed; and src/_ad_intg.hpp:220
// is rejected by g++ 13).  The reference holds no golden vectors for element
// residuals/Jacobians.  The AD layer of this oracle is pinned against the
// closed forms of the reference's own ex0.cpp:36-98 (tests/test_oracle_ad.py);
// the FE substrate (MFEM, un-vendored, version not pinned by the reference)
// is restated from MFEM's published algorithms and pinned by analytic
// identities (patch tests, manufactured solutions, finite differences).
// Element residual / Jacobian values: "parity unpinned" against a live MFEM.
//
// Every function cites the reference file:line it follows.  The algorithm is
// kept deliberately faithful to how the reference spends its time: dense
// per-element loops, physical shapes and J^-1 recomputed at every quadrature
// point, derivatives by seed-by-seed re-evaluation with nested duals.
//
// Build: make -C oracle   (g++ -O3 -march=native -shared -fPIC)

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace orc
{

// ---------------------------------------------------------------------------
// Dual numbers: semantics of mfem::future::dual as typedef'd at
// src/ad_native.hpp:42-49 (ADReal_t = dual<real,real>, AD2Real_t = dual<AD,AD>)
// ---------------------------------------------------------------------------
template <class V, class G> struct dual
{
   V value;
   G gradient;
};
typedef dual<double, double> D1;
typedef dual<D1, D1> D2;

inline double inner(double a) { return a; }
template <class V, class G> inline double inner(const dual<V, G> &a) { return inner(a.value); }

template <class T> struct Zero { static T get() { return T(); } };
template <> struct Zero<double> { static double get() { return 0.0; } };
template <class V, class G> struct Zero<dual<V, G>>
{
   static dual<V, G> get() { return dual<V, G> {Zero<V>::get(), Zero<G>::get()}; }
};
template <class T> inline T zero() { return Zero<T>::get(); }

template <class T> struct FromReal { static T get(double a) { return a; } };
template <class V, class G> struct FromReal<dual<V, G>>
{
   static dual<V, G> get(double a) { return dual<V, G> {FromReal<V>::get(a), Zero<G>::get()}; }
};

// + -
template <class V, class G> inline dual<V, G> operator+(const dual<V, G> &a, const dual<V, G> &b) { return {a.value + b.value, a.gradient + b.gradient}; }
template <class V, class G> inline dual<V, G> operator+(const dual<V, G> &a, double b) { return {a.value + b, a.gradient}; }
template <class V, class G> inline dual<V, G> operator+(double a, const dual<V, G> &b) { return {a + b.value, b.gradient}; }
template <class V, class G> inline dual<V, G> operator-(const dual<V, G> &a) { return {-a.value, -a.gradient}; }
template <class V, class G> inline dual<V, G> operator-(const dual<V, G> &a, const dual<V, G> &b) { return {a.value - b.value, a.gradient - b.gradient}; }
template <class V, class G> inline dual<V, G> operator-(const dual<V, G> &a, double b) { return {a.value - b, a.gradient}; }
template <class V, class G> inline dual<V, G> operator-(double a, const dual<V, G> &b) { return {a - b.value, -b.gradient}; }
// *
template <class V, class G> inline dual<V, G> operator*(const dual<V, G> &a, const dual<V, G> &b) { return {a.value * b.value, b.value * a.gradient + a.value * b.gradient}; }
template <class V, class G> inline dual<V, G> operator*(const dual<V, G> &a, double b) { return {a.value * b, a.gradient * b}; }
template <class V, class G> inline dual<V, G> operator*(double a, const dual<V, G> &b) { return {a * b.value, a * b.gradient}; }
// /
template <class V, class G> inline dual<V, G> operator/(const dual<V, G> &a, const dual<V, G> &b) { return {a.value / b.value, (a.gradient / b.value) - (a.value * b.gradient) / (b.value * b.value)}; }
template <class V, class G> inline dual<V, G> operator/(const dual<V, G> &a, double b) { return {a.value / b, a.gradient / b}; }
template <class V, class G> inline dual<V, G> operator/(double a, const dual<V, G> &b) { return {a / b.value, -(a * b.gradient) / (b.value * b.value)}; }
// compound
template <class V, class G, class O> inline dual<V, G> &operator+=(dual<V, G> &a, const O &b) { a = a + b; return a; }
template <class V, class G, class O> inline dual<V, G> &operator-=(dual<V, G> &a, const O &b) { a = a - b; return a; }
template <class V, class G, class O> inline dual<V, G> &operator*=(dual<V, G> &a, const O &b) { a = a * b; return a; }
template <class V, class G, class O> inline dual<V, G> &operator/=(dual<V, G> &a, const O &b) { a = a / b; return a; }
// comparisons act on values
template <class V, class G> inline bool operator>(const dual<V, G> &a, const dual<V, G> &b) { return inner(a) > inner(b); }
template <class V, class G> inline bool operator<(const dual<V, G> &a, const dual<V, G> &b) { return inner(a) < inner(b); }
template <class V, class G> inline bool operator>(const dual<V, G> &a, double b) { return inner(a) > b; }
template <class V, class G> inline bool operator<(const dual<V, G> &a, double b) { return inner(a) < b; }

// elementary functions
using std::cos;
using std::exp;
using std::log;
using std::pow;
using std::sin;
using std::sqrt;
template <class V, class G> inline dual<V, G> sqrt(const dual<V, G> &a) { return {sqrt(a.value), a.gradient / (2.0 * sqrt(a.value))}; }
template <class V, class G> inline dual<V, G> exp(const dual<V, G> &a) { return {exp(a.value), exp(a.value) * a.gradient}; }
template <class V, class G> inline dual<V, G> log(const dual<V, G> &a) { return {log(a.value), a.gradient / a.value}; }
template <class V, class G> inline dual<V, G> sin(const dual<V, G> &a) { return {sin(a.value), a.gradient * cos(a.value)}; }
template <class V, class G> inline dual<V, G> cos(const dual<V, G> &a) { return {cos(a.value), -a.gradient * sin(a.value)}; }
template <class V, class G> inline dual<V, G> pow(const dual<V, G> &a, double b) { return {pow(a.value, b), b * pow(a.value, b - 1.0) * a.gradient}; }

// dual-aware max/min, tie -> average: src/ad_native.hpp:695-749
inline double max(double a, double b) { return std::max(a, b); } // :27
inline double min(double a, double b) { return std::min(a, b); } // :35
template <class V, class G> inline dual<V, G> max(const dual<V, G> &a, const dual<V, G> &b)
{
   if (a > b) { return a; }
   else if (a < b) { return b; }
   else { return 0.5 * (a + b); }
}
template <class V, class G> inline dual<V, G> min(const dual<V, G> &a, const dual<V, G> &b)
{
   if (a < b) { return a; }
   else if (a > b) { return b; }
   else { return 0.5 * (a + b); }
}

// ---------------------------------------------------------------------------
// Vector view/owner with the subset of TAutoDiffVector / mfem::Vector the
// functional bodies use (SURVEY 8c "MFEM API subset").
// ---------------------------------------------------------------------------
template <class T> struct Vec
{
   T *d;
   int n;
   std::vector<T> own;
   Vec() : d(nullptr), n(0) {}
   Vec(T *p, int n_) : d(p), n(n_) {}
   explicit Vec(int n_) : d(nullptr), n(n_), own(n_, zero<T>()) { d = own.data(); }
   Vec(const Vec &o) : d(nullptr), n(o.n), own(o.d, o.d + o.n) { d = own.data(); } // deep copy
   Vec &operator=(const Vec &o)
   {
      own.assign(o.d, o.d + o.n); d = own.data(); n = o.n; return *this;
   }
   int Size() const { return n; }
   T *GetData() const { return d; }
   T &operator[](int i) const { return d[i]; }
   T &operator()(int i) const { return d[i]; }
   void SetDataAndSize(T *p, int n_) { d = p; n = n_; own.clear(); }
   template <class A> void Add(const A &a, const Vec &v) { for (int i = 0; i < n; i++) { d[i] += a * v[i]; } }
};
template <class T> inline T operator*(const Vec<T> &a, const Vec<T> &b)
{
   T s = zero<T>();
   for (int i = 0; i < a.n; i++) { s += a[i] * b[i]; }
   return s;
}

// ---------------------------------------------------------------------------
// Functional description (tree of nodes), shared with Python via ctypes
// ---------------------------------------------------------------------------
enum Kind
{
   K_EX0 = 1, K_MASS = 2, K_DIFFUSION = 3, K_DIFF = 4, K_ELASTICITY = 5,
   K_MINSURF = 6, K_OBSTACLE = 7, K_GRADOBSTACLE = 8, K_LAGRANGIAN = 9,
   K_AL = 10, K_PG = 11, K_LAMBDAPG = 12, K_SHANNON = 13, K_FERMIDIRAC = 14,
   K_HELLINGER = 15, K_SIMPLEX = 16, K_SIMP = 17, K_PARAMCOMPLIANCE = 18,
   K_EMPTY = 19, K_EX0VEC = 20, K_LOAD = 21
};

extern "C" struct orc_fn_t
{
   int kind;
   int n_input;
   int n_output;   // vector functions only
   int nparam;
   double param[24];
   int qoff;       // offset of this node's per-point parameters in qprm, or -1
   int nchild;
   int child[6];
   int iparam[8];
};

struct Ctx
{
   const orc_fn_t *nodes;
   const double *qprm; // Evaluator::val at this point (src/ad_native.cpp:120-179)
};

template <class T> T fn_eval(const Ctx &c, int id, const Vec<T> &x);

// src/ad_native.hpp:421-481 (DiffusionEnergy body)
template <class T> T diffusion(const orc_fn_t &f, const double *K, int Kdim, const Vec<T> &gradu)
{
   const int dim = gradu.Size();
   if (Kdim == 0) { return 0.5 * (gradu * gradu); }
   if (Kdim == 1) { return 0.5 * K[0] * (gradu * gradu); }
   if (Kdim == dim)
   {
      T result = zero<T>();
      for (int i = 0; i < dim; i++) { result += K[i] * gradu[i] * gradu[i]; }
      return 0.5 * result;
   }
   if (Kdim == dim * dim)
   {
      T result = zero<T>();
      for (int j = 0; j < dim; j++)
      {
         for (int i = 0; i < dim; i++) { result += K[i + dim * j] * gradu[i] * gradu[j]; }
      }
      return 0.5 * result;
   }
   fprintf(stderr, "oracle: DiffusionEnergy: bad K size\n");
   abort();
}

// src/ad_native.hpp:550-565 and src/mmto.hpp:173-188 (identical bodies)
template <class T> T elasticity(int dim, double lambda, double mu, const Vec<T> &gradu)
{
   T divnorm = zero<T>();
   for (int i = 0; i < dim; i++) { divnorm += gradu[i * dim + i]; }
   divnorm = divnorm * divnorm;
   T h1_norm = zero<T>();
   for (int i = 0; i < dim; i++)
   {
      for (int j = 0; j < dim; j++)
      {
         T symm = 0.5 * (gradu[i * dim + j] + gradu[j * dim + i]);
         h1_norm += symm * symm;
      }
   }
   return 0.5 * lambda * divnorm + mu * h1_norm;
}

template <class T> T fn_eval(const Ctx &c, int id, const Vec<T> &x)
{
   const orc_fn_t &f = c.nodes[id];
   const double *qp = (f.qoff >= 0 && c.qprm) ? c.qprm + f.qoff : nullptr;
   switch (f.kind)
   {
      case K_EX0: // ex0.cpp:20
         return sin(x(0)) * exp(x(1)) + pow(x(2), 3.0);
      case K_MASS: // src/ad_native.hpp:419
         return 0.5 * (x * x);
      case K_DIFFUSION:
      {
         const int Kdim = f.iparam[0];
         return diffusion<T>(f, qp ? qp : f.param, Kdim, x);
      }
      case K_DIFF: // src/ad_native.hpp:518-524
      {
         const double *target = qp ? qp : f.param;
         Vec<T> diff(x);
         for (int i = 0; i < f.n_input; i++) { diff[i] -= target[i]; }
         return fn_eval<T>(c, f.child[0], diff);
      }
      case K_ELASTICITY:
      {
         const double lambda = qp ? qp[0] : f.param[0];
         const double mu = qp ? qp[1] : f.param[1];
         return elasticity<T>(f.iparam[0], lambda, mu, x);
      }
      case K_PARAMCOMPLIANCE: // src/mmto.hpp:169-170: lambda, mu live in evaluator slots
         return elasticity<T>(f.iparam[0], qp[0], qp[1], x);
      case K_MINSURF: // ex2.cpp:17-23
      {
         const double eps = f.param[0];
         T h1_norm(x * x);
         return sqrt(h1_norm + 1.0) + eps * h1_norm;
      }
      case K_OBSTACLE: // ex4.cpp:18-27
      {
         T result = zero<T>();
         for (int i = 1; i < x.Size(); i++) { result += x[i] * x[i]; }
         return result * 0.5;
      }
      case K_GRADOBSTACLE: // ex5.cpp:18-21
         return x * x * 0.5;
      case K_EMPTY: // src/_dof_pg.hpp:14
         return zero<T>();
      case K_LOAD: // linear form (f, v): the gradient of sum_c f_c(x) u_c is MFEM's (Vector)DomainLFIntegrator(f) (ex4.cpp:145-148, ex3.cpp:64-67)
      {
         T result = x[0] * (qp ? qp[0] : f.param[0]);
         for (int c = 1; c < f.n_input; c++) { result += x[c] * (qp ? qp[c] : f.param[c]); }
         return result;
      }
      case K_LAGRANGIAN: // src/ad_native.hpp:607-618
      {
         const int nobj = c.nodes[f.child[0]].n_input;
         const int ncon = f.nchild - 1;
         const int eval_mode = f.iparam[0]; // -2 objective, -1 full, >=0 constraint
         const Vec<T> xx(x.GetData(), nobj);
         const Vec<T> lambda(x.GetData() + nobj, ncon);
         if (eval_mode >= 0) { return fn_eval<T>(c, f.child[1 + eval_mode], xx); }
         T result = fn_eval<T>(c, f.child[0], xx);
         if (eval_mode == -2) { return result; }
         for (int i = 0; i < ncon; i++) { result += fn_eval<T>(c, f.child[1 + i], xx) * lambda[i]; }
         return result;
      }
      case K_AL: // src/ad_native.hpp:670-690
      {
         const int ncon = f.nchild - 1;
         const int mode = f.iparam[0];
         const double penalty = f.param[0];
         const double *eq_rhs = f.param + 1;
         const double *lambda = f.param + 1 + ncon;
         auto evalAL = [&](int idx) -> T
         {
            T cx = fn_eval<T>(c, f.child[1 + idx], x) - eq_rhs[idx];
            if (mode >= 0) { return cx; }
            return cx * (lambda[idx] + penalty * 0.5 * cx);
         };
         if (mode >= 0) { return evalAL(mode); }
         T result = fn_eval<T>(c, f.child[0], x);
         if (mode == -2) { return result; }
         for (int i = 0; i < ncon; i++) { result += evalAL(i); }
         return result;
      }
      case K_PG: // src/pg.hpp:193-213
      {
         const orc_fn_t &fo = c.nodes[f.child[0]];
         const int nent = f.nchild - 1;
         const double alpha = f.param[0];
         const Vec<T> xx(x.GetData(), fo.n_input);
         Vec<T> psi;
         T cross_entropy = zero<T>();
         T dual_entropy_sum = zero<T>();
         int dual_idx = fo.n_input; // src/pg.hpp:104
         int koff = 0;
         for (int i = 0; i < nent; i++)
         {
            const int esz = c.nodes[f.child[1 + i]].n_input;
            psi.SetDataAndSize(x.GetData() + dual_idx, esz);
            const double *psi_k = qp + koff;
            for (int j = 0; j < esz; j++)
            {
               cross_entropy += xx[f.iparam[i] + j] * (psi[j] - psi_k[j]);
            }
            dual_entropy_sum += fn_eval<T>(c, f.child[1 + i], psi);
            dual_idx += esz;
            koff += esz;
         }
         return fn_eval<T>(c, f.child[0], xx) + (cross_entropy - dual_entropy_sum) / alpha;
      }
      case K_LAMBDAPG: // src/pg.hpp:220-242
      {
         const orc_fn_t &fo = c.nodes[f.child[0]];
         const int nent = f.nchild - 1;
         const double alpha = f.param[0];
         const Vec<T> xx(x.GetData(), fo.n_input);
         Vec<T> x_i, lambda;
         T cross_entropy = zero<T>();
         T dual_entropy_sum = zero<T>();
         int dual_idx = fo.n_input;
         int koff = 0;
         for (int i = 0; i < nent; i++)
         {
            const int esz = c.nodes[f.child[1 + i]].n_input;
            x_i.SetDataAndSize(x.GetData() + f.iparam[i], esz);
            lambda.SetDataAndSize(x.GetData() + dual_idx, esz);
            Vec<T> psi(esz);
            for (int j = 0; j < esz; j++) { psi[j] = FromReal<T>::get(qp[koff + j]); } // psi = psi_k
            psi.Add(alpha, lambda);
            cross_entropy += x_i * lambda;
            dual_entropy_sum += fn_eval<T>(c, f.child[1 + i], psi);
            dual_idx += esz;
            koff += esz;
         }
         return fn_eval<T>(c, f.child[0], xx) + cross_entropy - dual_entropy_sum / alpha;
      }
      case K_SHANNON: // src/pg.hpp:277
      {
         const double bound = qp ? qp[0] : f.param[0];
         const int sign = f.iparam[0];
         return sign * (exp(x[0] * sign)) + bound * x[0];
      }
      case K_FERMIDIRAC: // src/pg.hpp:289-321, incl. the slot swap
      {
         // ctor args (lower, upper) are added to evaluator slots 0, 1 (:294-295)
         const double slot0 = qp ? qp[0] : f.param[0]; // the 'lower' argument
         const double slot1 = qp ? qp[1] : f.param[1]; // the 'upper' argument
         const double upper_bound = slot0; // member bound to slot 0 (:291)
         const double lower_bound = slot1; // member bound to slot 1 (:292)
         const double shift = lower_bound;         // :305
         const double scale = upper_bound - shift; // :306
         T z = x[0] * scale;
         if (z > 0) { return z + log(1.0 + exp(-z)) + shift * x[0]; }
         else { return log(1.0 + exp(z)) + shift * x[0]; }
      }
      case K_HELLINGER: // src/pg.hpp:341
      {
         const double scale = qp ? qp[0] : f.param[0];
         return sqrt(1 + (x * x) * (scale * scale));
      }
      case K_SIMPLEX: // src/pg.hpp:364-375
      {
         const double scale = qp ? qp[0] : f.param[0];
         T maxval = x[0];
         for (int i = 1; i < x.Size(); i++) { maxval = max(maxval, x[i]); }
         T sum_exp = zero<T>();
         for (int i = 0; i < x.Size(); i++) { sum_exp += exp(x[i] - maxval); }
         return scale * (maxval + log(sum_exp));
      }
      case K_SIMP: // src/mmto.hpp:19-27
      {
         const double *E = f.param;
         const double p = f.param[f.n_input];
         T result = zero<T>();
         for (int i = 0; i < x.Size(); i++) { result += E[i] * pow(x[i], p); }
         return result;
      }
      default:
         fprintf(stderr, "oracle: unknown functional kind %d\n", f.kind);
         abort();
   }
}

// vector functions (AD_VEC_IMPL): ex0.cpp:29-33
template <class T> void vecfn_eval(const Ctx &c, int id, const Vec<T> &x, Vec<T> &result)
{
   const orc_fn_t &f = c.nodes[id];
   switch (f.kind)
   {
      case K_EX0VEC:
         result[0] = sin(x[0] * x[1]);
         result[1] = cos(x[0] * x[1] * x[2]);
         return;
      default:
         fprintf(stderr, "oracle: unknown vector functional kind %d\n", f.kind);
         abort();
   }
}

// src/ad_native.cpp:188-201
static void fn_gradient(const Ctx &c, int id, const double *x, double *J)
{
   const int n = c.nodes[id].n_input;
   std::vector<D1> xs(n);
   for (int i = 0; i < n; i++) { xs[i] = D1 {x[i], 0.0}; }
   Vec<D1> x_ad(xs.data(), n);
   for (int i = 0; i < n; i++)
   {
      x_ad[i].gradient = 1.0;
      D1 result = fn_eval<D1>(c, id, x_ad);
      J[i] = result.gradient;
      x_ad[i].gradient = 0.0;
   }
}

// src/ad_native.cpp:211-230 ; H is n x n column-major
static void fn_hessian(const Ctx &c, int id, const double *x, double *H)
{
   const int n = c.nodes[id].n_input;
   std::vector<D2> xs(n);
   for (int i = 0; i < n; i++) { xs[i] = D2 {D1 {x[i], 0.0}, D1 {0.0, 0.0}}; }
   Vec<D2> x_ad(xs.data(), n);
   for (int i = 0; i < n; i++)
   {
      x_ad[i].value.gradient = 1.0;
      for (int j = 0; j <= i; j++)
      {
         x_ad[j].gradient.value = 1.0;
         D2 result = fn_eval<D2>(c, id, x_ad);
         H[j + n * i] = result.gradient.gradient;
         H[i + n * j] = result.gradient.gradient;
         x_ad[j].gradient.value = 0.0;
      }
      x_ad[i].value.gradient = 0.0;
   }
}

static double fn_value(const Ctx &c, int id, const double *x)
{
   const int n = c.nodes[id].n_input;
   Vec<double> xv(const_cast<double *>(x), n);
   return fn_eval<double>(c, id, xv);
}

// ---------------------------------------------------------------------------
// FE substrate: restatement of the MFEM pieces the hot path calls
// (SURVEY 8c): Gauss-Legendre / Gauss-Lobatto points, 1-D Lagrange bases,
// tensor-product H1 (closed GLL) and L2 (open GL) elements, isoparametric
// transformation, tensor Gauss rules with x-fastest ordering.
// ---------------------------------------------------------------------------
static void gauss_legendre(int n, double *x, double *w) // on [0,1]
{
   for (int i = 0; i < (n + 1) / 2; i++)
   {
      double z = std::cos(M_PI * (i + 0.75) / (n + 0.5));
      double pp = 0.0;
      for (int it = 0; it < 100; it++)
      {
         double p1 = 1.0, p2 = 0.0;
         for (int j = 1; j <= n; j++)
         {
            double p3 = p2;
            p2 = p1;
            p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
         }
         pp = n * (z * p1 - p2) / (z * z - 1.0);
         double dz = p1 / pp;
         z -= dz;
         if (std::fabs(dz) < 1e-16) { break; }
      }
      // recompute pp at converged z
      {
         double p1 = 1.0, p2 = 0.0;
         for (int j = 1; j <= n; j++)
         {
            double p3 = p2;
            p2 = p1;
            p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
         }
         pp = n * (z * p1 - p2) / (z * z - 1.0);
      }
      x[i] = 0.5 * (1.0 - z);
      x[n - 1 - i] = 0.5 * (1.0 + z);
      w[i] = w[n - 1 - i] = 1.0 / ((1.0 - z * z) * pp * pp);
   }
}

static void gauss_lobatto(int n, double *x) // n points on [0,1], n >= 2
{
   x[0] = 0.0;
   x[n - 1] = 1.0;
   const int N = n - 1; // interior points are roots of P'_N
   for (int i = 1; i < n - 1; i++)
   {
      double z = -std::cos(M_PI * i / N); // Chebyshev-Lobatto initial guess
      for (int it = 0; it < 100; it++)
      {
         // P_N(z), P_{N-1}(z)
         double p1 = 1.0, p2 = 0.0;
         for (int j = 1; j <= N; j++)
         {
            double p3 = p2;
            p2 = p1;
            p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
         }
         // P'_N = N (z P_N - P_{N-1}) / (z^2-1); P''_N from Legendre ODE
         double dp = N * (z * p1 - p2) / (z * z - 1.0);
         double ddp = (2.0 * z * dp - N * (N + 1.0) * p1) / (1.0 - z * z);
         double dz = dp / ddp;
         z -= dz;
         if (std::fabs(dz) < 1e-16) { break; }
      }
      x[i] = 0.5 * (1.0 + z);
   }
   // symmetrise
   for (int i = 0; i < n / 2; i++)
   {
      double a = 0.5 * (x[i] + (1.0 - x[n - 1 - i]));
      x[i] = a;
      x[n - 1 - i] = 1.0 - a;
   }
   if (n % 2 == 1) { x[n / 2] = 0.5; }
}

// 1-D Lagrange basis on given nodes: values and derivatives at point t
static void lagrange_1d(int nn, const double *nodes, double t, double *b, double *db)
{
   for (int i = 0; i < nn; i++)
   {
      double v = 1.0, denom = 1.0;
      for (int j = 0; j < nn; j++)
      {
         if (j == i) { continue; }
         v *= (t - nodes[j]);
         denom *= (nodes[i] - nodes[j]);
      }
      double dv = 0.0;
      for (int k = 0; k < nn; k++)
      {
         if (k == i) { continue; }
         double prod = 1.0;
         for (int j = 0; j < nn; j++)
         {
            if (j == i || j == k) { continue; }
            prod *= (t - nodes[j]);
         }
         dv += prod;
      }
      b[i] = v / denom;
      db[i] = dv / denom;
   }
}

enum { BASIS_H1 = 0, BASIS_L2 = 1 };

// Triangle elements (H1_TriangleElement orders 1, 2; L2_TriangleElement order 0): nodal bases on the reference triangle
// (0,0), (1,0), (0,1); dofs: vertices, then edge midpoints in MFEM's edge order (0,1), (1,2), (2,0).
static int tri_dofs(int basis, int p) { return basis == BASIS_H1 ? (p == 1 ? 3 : (p == 2 ? 6 : -1)) : (p == 0 ? 1 : -1); }
static void tri_shape(int basis, int p, const double *ip, double *shape, double *dshape, int dof)
{
   // barycentric coordinates and their gradients
   const double L[3] = {1.0 - ip[0] - ip[1], ip[0], ip[1]};
   const double gx[3] = {-1.0, 1.0, 0.0}, gy[3] = {-1.0, 0.0, 1.0};
   if (basis != BASIS_H1)
   {
      if (shape) { shape[0] = 1.0; }
      if (dshape) { dshape[0] = 0.0; dshape[dof] = 0.0; }
      return;
   }
   for (int v = 0; v < 3; v++)
   {
      const double s = (p == 1) ? L[v] : L[v] * (2.0 * L[v] - 1.0);
      const double d = (p == 1) ? 1.0 : 4.0 * L[v] - 1.0; // ds/dL_v
      if (shape) { shape[v] = s; }
      if (dshape) { dshape[v] = d * gx[v]; dshape[v + dof] = d * gy[v]; }
   }
   if (p == 2)
   {
      for (int e = 0; e < 3; e++)
      {
         const int a = e, b = (e + 1) % 3; // edges (0,1), (1,2), (2,0)
         if (shape) { shape[3 + e] = 4.0 * L[a] * L[b]; }
         if (dshape)
         {
            dshape[3 + e] = 4.0 * (gx[a] * L[b] + L[a] * gx[b]);
            dshape[3 + e + dof] = 4.0 * (gy[a] * L[b] + L[a] * gy[b]);
         }
      }
   }
}

struct TensorElement // H1_{Segment,Quadrilateral,Hex}Element / L2_* with lexicographic dofs; or a triangle element
{
   int dim, p, nn, dof;
   int simplex = 0, basis_ = 0;
   std::vector<double> nodes;
   TensorElement(int dim_, int p_, int basis, int simplex_ = 0) : dim(dim_), p(p_), nn(p_ + 1), simplex(simplex_), basis_(basis)
   {
      if (simplex)
      {
         dof = tri_dofs(basis, p_);
         return;
      }
      nodes.resize(nn);
      if (basis == BASIS_H1)
      {
         if (nn == 1) { nodes[0] = 0.5; }
         else { gauss_lobatto(nn, nodes.data()); }
      }
      else
      {
         std::vector<double> w(nn);
         gauss_legendre(nn, nodes.data(), w.data());
      }
      dof = 1;
      for (int d = 0; d < dim; d++) { dof *= nn; }
   }
   // CalcShape: shape[dof]
   void CalcShape(const double *ip, double *shape) const
   {
      if (simplex) { tri_shape(basis_, p, ip, shape, nullptr, dof); return; }
      std::vector<double> b(3 * nn), db(3 * nn);
      for (int d = 0; d < dim; d++) { lagrange_1d(nn, nodes.data(), ip[d], &b[d * nn], &db[d * nn]); }
      for (int i = 0; i < dof; i++)
      {
         int r = i;
         double v = 1.0;
         for (int d = 0; d < dim; d++) { v *= b[d * nn + r % nn]; r /= nn; }
         shape[i] = v;
      }
   }
   // CalcDShape: dshape[dof x dim] column-major
   void CalcDShape(const double *ip, double *dshape) const
   {
      if (simplex) { tri_shape(basis_, p, ip, nullptr, dshape, dof); return; }
      std::vector<double> b(3 * nn), db(3 * nn);
      for (int d = 0; d < dim; d++) { lagrange_1d(nn, nodes.data(), ip[d], &b[d * nn], &db[d * nn]); }
      for (int i = 0; i < dof; i++)
      {
         int idx[3] = {0, 0, 0};
         int r = i;
         for (int d = 0; d < dim; d++) { idx[d] = r % nn; r /= nn; }
         for (int k = 0; k < dim; k++)
         {
            double v = 1.0;
            for (int d = 0; d < dim; d++) { v *= (d == k) ? db[d * nn + idx[d]] : b[d * nn + idx[d]]; }
            dshape[i + dof * k] = v;
         }
      }
   }
};

// IntRules.Get(Geometry::TRIANGLE, order) for orders 0 - 6, restated from the published symmetric rules MFEM's
// intrules.cpp lists (centroid / 3-point orbits (a, a, 1-2a) / 6-point orbits (a, b, 1-a-b)); weights sum to 1/2.
// Orbit tables: {kind (1, 3, 6), a, b, weight}
static const double TRI_RULES[7][4][4] = {
   {{1, 0, 0, 0.5}},
   {{1, 0, 0, 0.5}},
   {{3, 1.0 / 6.0, 0, 1.0 / 6.0}},
   {{1, 0, 0, -0.28125}, {3, 0.2, 0, 25.0 / 96.0}},
   {{3, 0.091576213509770743460, 0, 0.054975871827660933819}, {3, 0.44594849091596488632, 0, 0.11169079483900573285}},
   {{1, 0, 0, 0.1125}, {3, 0.10128650732345633880, 0, 0.062969590272413576298}, {3, 0.47014206410511508977, 0, 0.066197076394253090369}},
   {{3, 0.063089014491502228340, 0, 0.025422453185103408460}, {3, 0.24928674517091042129, 0, 0.058393137863189683013},
    {6, 0.053145049844816947353, 0.31035245103378440542, 0.041425537809186787597}}};

struct IntRule // IntRules.Get(SEGMENT/SQUARE/CUBE, order): tensor Gauss-Legendre; TRIANGLE: the tables above
{
   int dim, n1, np;
   std::vector<double> pts, w; // pts[np*dim]
   IntRule(int dim_, int order, int simplex = 0) : dim(dim_)
   {
      if (simplex)
      {
         n1 = 0;
         const int o = std::min(std::max(order, 0), 6);
         for (int k = 0; k < 4; k++)
         {
            const double *r = TRI_RULES[o][k];
            const int kind = (int)r[0];
            if (kind == 0) { break; }
            const double a = r[1], b = r[2], wt = r[3];
            if (kind == 1) { pts.insert(pts.end(), {1.0 / 3.0, 1.0 / 3.0}); w.push_back(wt); }
            else if (kind == 3)
            {
               const double c = 1.0 - 2.0 * a;
               pts.insert(pts.end(), {a, a, a, c, c, a});
               w.insert(w.end(), {wt, wt, wt});
            }
            else
            {
               const double c = 1.0 - a - b;
               pts.insert(pts.end(), {a, b, b, a, a, c, c, a, b, c, c, b});
               w.insert(w.end(), {wt, wt, wt, wt, wt, wt});
            }
         }
         np = (int)w.size();
         return;
      }
      const int real_order = order | 1; // GetSegmentRealOrder
      n1 = real_order / 2 + 1;
      std::vector<double> x1(n1), w1(n1);
      gauss_legendre(n1, x1.data(), w1.data());
      np = 1;
      for (int d = 0; d < dim; d++) { np *= n1; }
      pts.resize(np * dim);
      w.resize(np);
      for (int q = 0; q < np; q++) // x fastest
      {
         int r = q;
         double ww = 1.0;
         for (int d = 0; d < dim; d++)
         {
            pts[q * dim + d] = x1[r % n1];
            ww *= w1[r % n1];
            r /= n1;
         }
         w[q] = ww;
      }
   }
};

// dense helpers (column-major), following the mfem::DenseMatrix kernels used
static void invert_small(int n, const double *J, double *Ji, double *det)
{
   if (n == 1) { *det = J[0]; Ji[0] = 1.0 / J[0]; }
   else if (n == 2)
   {
      double d = J[0] * J[3] - J[1] * J[2];
      *det = d;
      double t = 1.0 / d;
      Ji[0] = J[3] * t; Ji[1] = -J[1] * t; Ji[2] = -J[2] * t; Ji[3] = J[0] * t;
   }
   else
   {
      const double *a = J; // a[i + 3*j]
      double c00 = a[4] * a[8] - a[5] * a[7];
      double c01 = a[5] * a[6] - a[3] * a[8];
      double c02 = a[3] * a[7] - a[4] * a[6];
      double d = a[0] * c00 + a[1] * c01 + a[2] * c02;
      *det = d;
      double t = 1.0 / d;
      // inverse(i,j) = cof(j,i)/det
      Ji[0] = c00 * t;
      Ji[1] = (a[2] * a[7] - a[1] * a[8]) * t;
      Ji[2] = (a[1] * a[5] - a[2] * a[4]) * t;
      Ji[3] = c01 * t;
      Ji[4] = (a[0] * a[8] - a[2] * a[6]) * t;
      Ji[5] = (a[2] * a[3] - a[0] * a[5]) * t;
      Ji[6] = c02 * t;
      Ji[7] = (a[1] * a[6] - a[0] * a[7]) * t;
      Ji[8] = (a[0] * a[4] - a[1] * a[3]) * t;
   }
}

// ---------------------------------------------------------------------------
// Form description
// ---------------------------------------------------------------------------
enum // ADEval flags, src/_ad_intg.hpp:24-36
{
   EV_QVALUE = 1 << 0, EV_VALUE = 1 << 1, EV_GRAD = 1 << 2, EV_DIV = 1 << 3,
   EV_CURL = 1 << 4, EV_HESSIAN = 1 << 5, EV_VECTOR = 1 << 6, EV_VECFE = 1 << 7
};
enum { ORD_BYNODES = 0, ORD_BYVDIM = 1 };
enum { PRM_CONST = 0, PRM_GF = 1, PRM_QF = 2, PRM_GF_GRAD = 3 };

extern "C" struct orc_space_t
{
   int basis, order, vdim, mode, ordering, ndofs; // ndofs = scalar dofs
   const int *e2l; // [ne * dof_el], lexicographic local order, scalar dof ids
};
extern "C" struct orc_mesh_t
{
   int dim, ne, geom_order, nnodes;
   const int *e2n;       // [ne * (geom_order+1)^dim], lexicographic
   const double *coords; // [nnodes * dim], xyzxyz
};
extern "C" struct orc_param_t
{
   int type, size;
   orc_space_t space;  // PRM_GF / PRM_GF_GRAD: the GridFunction's space
   const double *data; // CONST: values; GF: dof vector; QF: [ne*nq*size]
};
extern "C" struct orc_form_t
{
   orc_mesh_t mesh;
   int nspaces;
   const orc_space_t *spaces;
   const orc_fn_t *fn;
   int root;
   int quad_order; // <0: default 2*max_order+2 (src/_ad_intg.hpp:99-105, :298-313)
   int nparams;
   const orc_param_t *params;
   int block;      // 0: ADNonlinearFormIntegrator<mode>; 1: ADBlockNonlinearFormIntegrator<modes...>
   int ness;
   const int *ess; // essential dofs, global (concatenated) numbering
};

static int shapedim_of(int mode, int dim)
{
   // src/ad_intg.hpp:76-87
   int sd = 0;
   if (mode & EV_QVALUE) { sd += 1; }
   if (mode & EV_VALUE) { sd += 1; }
   if (mode & EV_GRAD) { sd += dim; }
   if (mode & EV_DIV) { sd += 1; }
   return sd;
}

struct ElemCtx // per-form scratch: elements, rule, offsets
{
   const orc_form_t &F;
   int dim;
   TensorElement geom;
   std::vector<TensorElement> els;
   std::vector<TensorElement> pels;
   IntRule ir;
   std::vector<int> dof, vdim, sd, xoff, voff; // per space
   int n_input, nvd_total;
   std::vector<int> goff; // global block offsets
   std::vector<int> poff; // per-point parameter offsets
   int nprm;
   static int qorder(const orc_form_t &F)
   {
      if (F.quad_order >= 0) { return F.quad_order; }
      int order = 0;
      for (int s = 0; s < F.nspaces; s++) { order = std::max(order, F.spaces[s].order); }
      return 2 * order + 2;
   }
   ElemCtx(const orc_form_t &F_)
      : F(F_), dim(F_.mesh.dim), geom(F_.mesh.dim, F_.mesh.geom_order < 0 ? 1 : F_.mesh.geom_order, BASIS_H1, F_.mesh.geom_order < 0),
        ir(F_.mesh.dim, qorder(F_), F_.mesh.geom_order < 0)
   {
      const int simplex = F.mesh.geom_order < 0 ? 1 : 0; // mesh.geom_order = -1: triangles (3 vertices, affine map)
      n_input = 0;
      nvd_total = 0;
      goff.push_back(0);
      for (int s = 0; s < F.nspaces; s++)
      {
         const orc_space_t &S = F.spaces[s];
         els.emplace_back(dim, S.order, S.basis, simplex);
         dof.push_back(els.back().dof);
         vdim.push_back(S.vdim);
         sd.push_back(shapedim_of(S.mode, dim));
         xoff.push_back(n_input);
         voff.push_back(nvd_total);
         n_input += sd.back() * S.vdim;
         nvd_total += dof.back() * S.vdim;
         goff.push_back(goff.back() + S.ndofs * S.vdim);
      }
      nprm = 0;
      for (int i = 0; i < F.nparams; i++)
      {
         poff.push_back(nprm);
         nprm += F.params[i].size;
         const orc_param_t &P = F.params[i];
         if (P.type == PRM_GF || P.type == PRM_GF_GRAD) { pels.emplace_back(dim, P.space.order, P.space.basis, simplex); }
         else { pels.emplace_back(dim, 0, BASIS_L2, simplex); }
      }
   }
   int vdof(int s, int e, int i, int c) const // global index in the concatenated vector
   {
      const orc_space_t &S = F.spaces[s];
      const int d = S.e2l[(size_t)e * dof[s] + i];
      return goff[s] + (S.ordering == ORD_BYNODES ? d + S.ndofs * c : d * S.vdim + c);
   }
};

struct PointGeom
{
   double J[9], Ji[9], detJ;
};

// ElementTransformation::SetIntPoint + Jacobian/Weight/InverseJacobian
static void eval_geom(const ElemCtx &C, int e, const double *ip, PointGeom &g)
{
   const int dim = C.dim, nn = C.geom.dof;
   std::vector<double> dsh(nn * dim);
   C.geom.CalcDShape(ip, dsh.data());
   for (int i = 0; i < dim * dim; i++) { g.J[i] = 0.0; }
   for (int k = 0; k < nn; k++)
   {
      const double *X = C.F.mesh.coords + (size_t)C.F.mesh.e2n[(size_t)e * nn + k] * dim;
      for (int j = 0; j < dim; j++)
      {
         for (int i = 0; i < dim; i++) { g.J[i + dim * j] += X[i] * dsh[k + nn * j]; }
      }
   }
   invert_small(dim, g.J, g.Ji, &g.detJ);
}

// CalcInputShapes: src/ad_intg.hpp:118-154 (single) / :424-466 (block)
// allshapes[dof x shapedim] column-major
static void calc_input_shapes(const ElemCtx &C, int s, const double *ip, int ipindex, const PointGeom &g, double *allshapes)
{
   const int dim = C.dim, dof = C.dof[s], mode = C.F.spaces[s].mode;
   int col = 0;
   if (mode & EV_QVALUE)
   {
      for (int i = 0; i < dof; i++) { allshapes[i + dof * col] = 0.0; }
      allshapes[ipindex + dof * col] = 1.0; // :127
      col++;
   }
   if (mode & EV_VALUE)
   {
      C.els[s].CalcShape(ip, allshapes + dof * col); // CalcPhysShape == CalcShape (MapType VALUE)
      col++;
   }
   int gcol = col;
   if (mode & EV_GRAD)
   {
      std::vector<double> dsh(dof * dim);
      C.els[s].CalcDShape(ip, dsh.data());
      // CalcPhysDShape: gshape = dshape * J^-1
      for (int j = 0; j < dim; j++)
      {
         for (int i = 0; i < dof; i++)
         {
            double v = 0.0;
            for (int k = 0; k < dim; k++) { v += dsh[i + dof * k] * g.Ji[k + dim * j]; }
            allshapes[i + dof * (col + j)] = v;
         }
      }
      col += dim;
   }
   if (mode & EV_DIV)
   {
      // only DIV together with GRAD is defined: row sums of gshape (:142-145)
      for (int i = 0; i < dof; i++)
      {
         double v = 0.0;
         for (int j = 0; j < dim; j++) { v += allshapes[i + dof * (gcol + j)]; }
         allshapes[i + dof * col] = v;
      }
      col++;
   }
}

// Evaluator::Eval at (element, point): src/ad_native.cpp:120-179
static void eval_params(const ElemCtx &C, int e, int q, const double *ip, const PointGeom &g, double *qprm)
{
   const int dim = C.dim;
   for (int i = 0; i < C.F.nparams; i++)
   {
      const orc_param_t &P = C.F.params[i];
      double *v = qprm + C.poff[i];
      if (P.type == PRM_CONST)
      {
         for (int k = 0; k < P.size; k++) { v[k] = P.data[k]; }
      }
      else if (P.type == PRM_QF) // QuadratureFunction::GetValues(ElementNo, ip.index)
      {
         for (int k = 0; k < P.size; k++) { v[k] = P.data[((size_t)e * C.ir.np + q) * P.size + k]; }
      }
      else if (P.type == PRM_GF) // GridFunction::GetVectorValue
      {
         const TensorElement &el = C.pels[i];
         std::vector<double> sh(el.dof);
         el.CalcShape(ip, sh.data());
         for (int c = 0; c < P.space.vdim; c++)
         {
            double s = 0.0;
            for (int k = 0; k < el.dof; k++)
            {
               const int d = P.space.e2l[(size_t)e * el.dof + k];
               const int gd = (P.space.ordering == ORD_BYNODES) ? d + P.space.ndofs * c : d * P.space.vdim + c;
               s += sh[k] * P.data[gd];
            }
            v[c] = s;
         }
      }
      else if (P.type == PRM_GF_GRAD) // GridFunction::GetVectorGradient -> DenseMatrix [vdim x sdim] col-major (src/tools.hpp:20-33)
      {
         const TensorElement &el = C.pels[i];
         std::vector<double> dsh(el.dof * dim), gsh(el.dof * dim);
         el.CalcDShape(ip, dsh.data());
         for (int j = 0; j < dim; j++)
         {
            for (int k = 0; k < el.dof; k++)
            {
               double t = 0.0;
               for (int m = 0; m < dim; m++) { t += dsh[k + el.dof * m] * g.Ji[m + dim * j]; }
               gsh[k + el.dof * j] = t;
            }
         }
         const int vd = P.space.vdim;
         for (int j = 0; j < dim; j++)
         {
            for (int c = 0; c < vd; c++)
            {
               double s = 0.0;
               for (int k = 0; k < el.dof; k++)
               {
                  const int d = P.space.e2l[(size_t)e * el.dof + k];
                  const int gd = (P.space.ordering == ORD_BYNODES) ? d + P.space.ndofs * c : d * vd + c;
                  s += gsh[k + el.dof * j] * P.data[gd];
               }
               v[c + vd * j] = s;
            }
         }
      }
   }
}

// Gather element dofs: elfun_s[dof x vdim] (byNODES inside the element)
static void gather(const ElemCtx &C, int e, const double *x, std::vector<std::vector<double>> &elfun)
{
   elfun.resize(C.F.nspaces);
   for (int s = 0; s < C.F.nspaces; s++)
   {
      elfun[s].resize(C.dof[s] * C.vdim[s]);
      for (int c = 0; c < C.vdim[s]; c++)
      {
         for (int i = 0; i < C.dof[s]; i++) { elfun[s][i + C.dof[s] * c] = x[C.vdof(s, e, i, c)]; }
      }
   }
}

// x_s = allshapes_s^T elfun_s  (MultAtB for VECTOR: xmat[shapedim x vdim])
static void interp_inputs(const ElemCtx &C, const std::vector<std::vector<double>> &allshapes,
                          const std::vector<std::vector<double>> &elfun, double *x)
{
   for (int s = 0; s < C.F.nspaces; s++)
   {
      const int dof = C.dof[s], sd = C.sd[s];
      for (int c = 0; c < C.vdim[s]; c++)
      {
         for (int k = 0; k < sd; k++)
         {
            double v = 0.0;
            for (int i = 0; i < dof; i++) { v += allshapes[s][i + dof * k] * elfun[s][i + dof * c]; }
            x[C.xoff[s] + k + sd * c] = v;
         }
      }
   }
}

// GetElementEnergy: src/ad_intg.hpp:157-199 / :469-530
static double element_energy(const ElemCtx &C, int e, const double *xg)
{
   std::vector<std::vector<double>> elfun, allshapes(C.F.nspaces);
   gather(C, e, xg, elfun);
   for (int s = 0; s < C.F.nspaces; s++) { allshapes[s].resize(C.dof[s] * C.sd[s]); }
   std::vector<double> x(C.n_input), qprm(std::max(C.nprm, 1));
   double energy = 0.0;
   for (int q = 0; q < C.ir.np; q++)
   {
      const double *ip = &C.ir.pts[q * C.dim];
      PointGeom g;
      eval_geom(C, e, ip, g);
      for (int s = 0; s < C.F.nspaces; s++) { calc_input_shapes(C, s, ip, q, g, allshapes[s].data()); }
      interp_inputs(C, allshapes, elfun, x.data());
      eval_params(C, e, q, ip, g, qprm.data());
      Ctx c {C.F.fn, qprm.data()};
      energy += fn_value(c, C.F.root, x.data()) * g.detJ * C.ir.w[q];
   }
   return energy;
}

// AssembleElementVector: src/ad_intg.hpp:202-257 (single) / :533-619 (block)
// elvect: concatenation over spaces of [dof x vdim]
static void element_vector(const ElemCtx &C, int e, const double *xg, double *elvect)
{
   std::vector<std::vector<double>> elfun, allshapes(C.F.nspaces);
   gather(C, e, xg, elfun);
   for (int s = 0; s < C.F.nspaces; s++) { allshapes[s].resize(C.dof[s] * C.sd[s]); }
   std::vector<double> x(C.n_input), jac(C.n_input), qprm(std::max(C.nprm, 1));
   for (int i = 0; i < C.nvd_total; i++) { elvect[i] = 0.0; }
   for (int q = 0; q < C.ir.np; q++)
   {
      const double *ip = &C.ir.pts[q * C.dim];
      PointGeom g;
      eval_geom(C, e, ip, g);
      const double w = C.ir.w[q] * g.detJ;
      for (int s = 0; s < C.F.nspaces; s++) { calc_input_shapes(C, s, ip, q, g, allshapes[s].data()); }
      interp_inputs(C, allshapes, elfun, x.data());
      eval_params(C, e, q, ip, g, qprm.data());
      Ctx c {C.F.fn, qprm.data()};
      fn_gradient(c, C.F.root, x.data(), jac.data());
      for (int i = 0; i < C.n_input; i++) { jac[i] *= w; }
      for (int s = 0; s < C.F.nspaces; s++)
      {
         const int dof = C.dof[s], sd = C.sd[s];
         // AddMult(allshapes, jacMat, elvectmat)
         for (int cc = 0; cc < C.vdim[s]; cc++)
         {
            for (int k = 0; k < sd; k++)
            {
               const double jv = jac[C.xoff[s] + k + sd * cc];
               for (int i = 0; i < dof; i++) { elvect[C.voff[s] + i + dof * cc] += allshapes[s][i + dof * k] * jv; }
            }
         }
      }
   }
}

// AssembleElementGrad.  elmat: [nvd_total x nvd_total] column-major over the
// concatenated element vector (block (test,trial) at (voff[test], voff[trial])).
static void element_grad(const ElemCtx &C, int e, const double *xg, double *elmat)
{
   const int N = C.nvd_total, n = C.n_input;
   std::vector<std::vector<double>> elfun, allshapes(C.F.nspaces);
   gather(C, e, xg, elfun);
   for (int s = 0; s < C.F.nspaces; s++) { allshapes[s].resize(C.dof[s] * C.sd[s]); }
   std::vector<double> x(n), H(n * n), qprm(std::max(C.nprm, 1));
   for (size_t i = 0; i < (size_t)N * N; i++) { elmat[i] = 0.0; }
   for (int q = 0; q < C.ir.np; q++)
   {
      const double *ip = &C.ir.pts[q * C.dim];
      PointGeom g;
      eval_geom(C, e, ip, g);
      const double w = C.ir.w[q] * g.detJ;
      for (int s = 0; s < C.F.nspaces; s++) { calc_input_shapes(C, s, ip, q, g, allshapes[s].data()); }
      interp_inputs(C, allshapes, elfun, x.data());
      eval_params(C, e, q, ip, g, qprm.data());
      Ctx c {C.F.fn, qprm.data()};
      fn_hessian(c, C.F.root, x.data(), H.data());
      for (int i = 0; i < n * n; i++) { H[i] *= w; }

      if (!C.F.block)
      {
         // ---- ADNonlinearFormIntegrator<mode>: src/ad_intg.hpp:310-332 ----
         const int dof = C.dof[0], sd = C.sd[0], vd = C.vdim[0];
         const double *B = allshapes[0].data();
         if (C.F.spaces[0].mode & EV_VECTOR)
         {
            // Hs = H viewed [sd x (vd*sd*vd)] (:292); Hx = allshapes * Hs (:312)
            const int ncol = vd * sd * vd;
            std::vector<double> Hx((size_t)dof * ncol), part((size_t)dof * dof);
            for (int k = 0; k < ncol; k++)
            {
               for (int i = 0; i < dof; i++)
               {
                  double v = 0.0;
                  for (int s1 = 0; s1 < sd; s1++) { v += B[i + dof * s1] * H[s1 + sd * k]; }
                  Hx[i + (size_t)dof * k] = v;
               }
            }
            const int nel = sd * dof;
            for (int cc = 0; cc < vd; cc++)
            {
               for (int r = 0; r <= cc; r++)
               {
                  // Hxsub = contiguous window [dof x sd] at (c*vdim + r)*nel (:318)
                  const double *Hxsub = Hx.data() + (size_t)(cc * vd + r) * nel;
                  // MultABt(allshapes, Hxsub, partelmat) (:319)
                  for (int j = 0; j < dof; j++)
                  {
                     for (int i = 0; i < dof; i++)
                     {
                        double v = 0.0;
                        for (int t = 0; t < sd; t++) { v += B[i + dof * t] * Hxsub[j + dof * t]; }
                        part[i + (size_t)dof * j] = v;
                     }
                  }
                  for (int j = 0; j < dof; j++)
                  {
                     for (int i = 0; i < dof; i++)
                     {
                        elmat[(cc * dof + i) + (size_t)N * (r * dof + j)] += part[i + (size_t)dof * j]; // :320
                        if (cc != r) { elmat[(r * dof + i) + (size_t)N * (cc * dof + j)] += part[i + (size_t)dof * j]; } // :323
                     }
                  }
               }
            }
         }
         else
         {
            // Hx = allshapes * H ; elmat += allshapes * Hx^T (:330-331)
            std::vector<double> Hx((size_t)dof * n);
            for (int k = 0; k < n; k++)
            {
               for (int i = 0; i < dof; i++)
               {
                  double v = 0.0;
                  for (int s1 = 0; s1 < sd; s1++) { v += B[i + dof * s1] * H[s1 + n * k]; }
                  Hx[i + (size_t)dof * k] = v;
               }
            }
            for (int j = 0; j < dof; j++)
            {
               for (int i = 0; i < dof; i++)
               {
                  double v = 0.0;
                  for (int k = 0; k < n; k++) { v += B[i + dof * k] * Hx[j + (size_t)dof * k]; }
                  elmat[i + (size_t)N * j] += v;
               }
            }
         }
      }
      else
      {
         // ---- ADBlockNonlinearFormIntegrator: src/ad_intg.hpp:700-727 ----
         for (int tr = 0; tr < C.F.nspaces; tr++)
         {
            for (int ts = 0; ts < C.F.nspaces; ts++)
            {
               const int tr_vdim = C.vdim[tr], ts_vdim = C.vdim[ts];
               const int sdt = C.sd[ts], sdr = C.sd[tr];
               const int doft = C.dof[ts], dofr = C.dof[tr];
               const int nr = sdt * ts_vdim, nc = sdr * tr_vdim;
               // Hsub = H[test rows, trial cols] (:708), column-major nr x nc
               std::vector<double> Hsub((size_t)nr * nc);
               for (int b = 0; b < nc; b++)
               {
                  for (int a = 0; a < nr; a++) { Hsub[a + (size_t)nr * b] = H[(C.xoff[ts] + a) + n * (C.xoff[tr] + b)]; }
               }
               // reinterpret [sdt x (ts_vdim*tr_vdim*sdr)] (:711); Hx = B_test * Hsub (:713)
               const int ncol = ts_vdim * tr_vdim * sdr;
               std::vector<double> Hx((size_t)doft * ncol);
               const double *Bt = allshapes[ts].data();
               for (int k = 0; k < ncol; k++)
               {
                  for (int i = 0; i < doft; i++)
                  {
                     double v = 0.0;
                     for (int s1 = 0; s1 < sdt; s1++) { v += Bt[i + doft * s1] * Hsub[s1 + (size_t)sdt * k]; }
                     Hx[i + (size_t)doft * k] = v;
                  }
               }
               // reinterpret Hx as [(doft*ts_vdim) x (tr_vdim*sdr)] (:714)
               const int h = doft * ts_vdim, w_ = sdr, wout = dofr;
               const double *Br = allshapes[tr].data();
               for (int d = 0; d < tr_vdim; d++)
               {
                  const double *Hxsub = Hx.data() + (size_t)d * (w_ * h); // [h x w_]
                  // partelmat = elmat(test,trial) + d*wout*h, [h x wout]; MyAddMultABt(Hxsub, B_trial) (:723), k-outer loop (:36-50)
                  for (int k = 0; k < w_; k++)
                  {
                     for (int j = 0; j < wout; j++)
                     {
                        const double bjk = Br[j + dofr * k];
                        for (int i = 0; i < h; i++)
                        {
                           elmat[(C.voff[ts] + i) + (size_t)N * (C.voff[tr] + d * wout + j)] += Hxsub[i + (size_t)h * k] * bjk;
                        }
                     }
                  }
               }
            }
         }
      }
   }
}

// element vdofs in the concatenated global numbering, order = element vector order
static void element_vdofs(const ElemCtx &C, int e, std::vector<int> &vd)
{
   vd.resize(C.nvd_total);
   for (int s = 0; s < C.F.nspaces; s++)
   {
      for (int c = 0; c < C.vdim[s]; c++)
      {
         for (int i = 0; i < C.dof[s]; i++) { vd[C.voff[s] + i + C.dof[s] * c] = C.vdof(s, e, i, c); }
      }
   }
}

} // namespace orc

using namespace orc;

// ---------------------------------------------------------------------------
// C entry points (ctypes)
// ---------------------------------------------------------------------------
extern "C"
{

   double orc_fn_value(const orc_fn_t *nodes, int root, const double *x, const double *qprm)
   {
      Ctx c {nodes, qprm};
      return fn_value(c, root, x);
   }
   void orc_fn_gradient(const orc_fn_t *nodes, int root, const double *x, const double *qprm, double *J)
   {
      Ctx c {nodes, qprm};
      fn_gradient(c, root, x, J);
   }
   void orc_fn_hessian(const orc_fn_t *nodes, int root, const double *x, const double *qprm, double *H)
   {
      Ctx c {nodes, qprm};
      fn_hessian(c, root, x, H);
   }
   // ADVectorFunction::Gradient: src/ad_native.cpp:232-250 ; J[n_out x n_in] col-major
   void orc_vecfn_gradient(const orc_fn_t *nodes, int root, const double *x, double *J)
   {
      Ctx c {nodes, nullptr};
      const int n = nodes[root].n_input, m = nodes[root].n_output;
      std::vector<D1> xs(n), fs(m);
      for (int i = 0; i < n; i++) { xs[i] = D1 {x[i], 0.0}; }
      Vec<D1> x_ad(xs.data(), n), Fx(fs.data(), m);
      for (int i = 0; i < n; i++)
      {
         x_ad[i].gradient = 1.0;
         for (int j = 0; j < m; j++) { Fx[j] = D1 {0.0, 0.0}; }
         vecfn_eval<D1>(c, root, x_ad, Fx);
         for (int j = 0; j < m; j++) { J[j + m * i] = Fx[j].gradient; }
         x_ad[i].gradient = 0.0;
      }
   }
   // ADVectorFunction::Hessian: src/ad_native.cpp:252-276 ; H[n_in x n_in x n_out]
   void orc_vecfn_hessian(const orc_fn_t *nodes, int root, const double *x, double *H)
   {
      Ctx c {nodes, nullptr};
      const int n = nodes[root].n_input, m = nodes[root].n_output;
      std::vector<D2> xs(n), fs(m);
      for (int i = 0; i < n; i++) { xs[i] = D2 {D1 {x[i], 0.0}, D1 {0.0, 0.0}}; }
      Vec<D2> x_ad(xs.data(), n), Fx(fs.data(), m);
      for (int i = 0; i < n; i++)
      {
         x_ad[i].value.gradient = 1.0;
         for (int j = 0; j <= i; j++)
         {
            x_ad[j].gradient.value = 1.0;
            for (int k = 0; k < m; k++) { Fx[k] = zero<D2>(); }
            vecfn_eval<D2>(c, root, x_ad, Fx);
            for (int k = 0; k < m; k++)
            {
               H[j + n * i + n * n * k] = Fx[k].gradient.gradient;
               H[i + n * j + n * n * k] = Fx[k].gradient.gradient;
            }
            x_ad[j].gradient.value = 0.0;
         }
         x_ad[i].value.gradient = 0.0;
      }
   }
   void orc_vecfn_value(const orc_fn_t *nodes, int root, const double *x, double *F)
   {
      Ctx c {nodes, nullptr};
      Vec<double> xv(const_cast<double *>(x), nodes[root].n_input), Fv(F, nodes[root].n_output);
      vecfn_eval<double>(c, root, xv, Fv);
   }

   // PGStepSizeRule::Get: src/pg.cpp:34-54
   double orc_pg_step(int rule_type, double alpha0, double max_alpha, double ratio, double ratio2, int iter)
   {
      double alpha = alpha0;
      switch (rule_type)
      {
         case 0: break;
         case 1: alpha *= std::pow(iter + 1, ratio); break;
         case 2: alpha *= std::pow(ratio, iter); break;
         case 3: alpha *= std::pow(ratio, std::pow(ratio2, iter)); break;
         default: break;
      }
      return std::min(alpha, max_alpha);
   }

   // 1-D tables for cross-checks
   void orc_gauss_legendre(int n, double *x, double *w) { gauss_legendre(n, x, w); }
   void orc_gauss_lobatto(int n, double *x) { gauss_lobatto(n, x); }
   int orc_rule_npts_1d(int order) { return (order | 1) / 2 + 1; }

   /// reference points [nq*dim] and weights [nq] of the form's integration rule
   int orc_form_rule(const orc_form_t *F, double *pts, double *w)
   {
      ElemCtx C(*F);
      for (size_t i = 0; i < C.ir.pts.size(); i++) { pts[i] = C.ir.pts[i]; }
      for (size_t i = 0; i < C.ir.w.size(); i++) { w[i] = C.ir.w[i]; }
      return 0;
   }
   int orc_form_sizes(const orc_form_t *F, int *n_input, int *nvd_el, int *nq, int *ndof_total)
   {
      ElemCtx C(*F);
      *n_input = C.n_input;
      *nvd_el = C.nvd_total;
      *nq = C.ir.np;
      *ndof_total = C.goff.back();
      return 0;
   }

   // element-level entry points
   double orc_element_energy(const orc_form_t *F, int e, const double *x)
   {
      ElemCtx C(*F);
      return element_energy(C, e, x);
   }
   void orc_element_vector(const orc_form_t *F, int e, const double *x, double *elvect)
   {
      ElemCtx C(*F);
      element_vector(C, e, x, elvect);
   }
   void orc_element_grad(const orc_form_t *F, int e, const double *x, double *elmat)
   {
      ElemCtx C(*F);
      element_grad(C, e, x, elmat);
   }

   // NonlinearForm::GetEnergy [MFEM-upstream]: sum of element energies
   double orc_form_energy(const orc_form_t *F, const double *x)
   {
      ElemCtx C(*F);
      double en = 0.0;
      for (int e = 0; e < F->mesh.ne; e++) { en += element_energy(C, e, x); }
      return en;
   }

   // NonlinearForm::Mult / BlockNonlinearForm::MultBlocked [MFEM-upstream]:
   // gather, AssembleElementVector, AddElementVector in element order, y[ess]=0.
   // [e0,e1) restricts the element loop (CPU-baseline sampling); y must be zeroed by the caller if e0>0.
   void orc_form_mult_range(const orc_form_t *F, const double *x, double *y, int e0, int e1, int zero_y)
   {
      ElemCtx C(*F);
      if (zero_y) { for (int i = 0; i < C.goff.back(); i++) { y[i] = 0.0; } }
      std::vector<double> elvect(C.nvd_total);
      std::vector<int> vd;
      for (int e = e0; e < e1; e++)
      {
         element_vector(C, e, x, elvect.data());
         element_vdofs(C, e, vd);
         for (int i = 0; i < C.nvd_total; i++) { y[vd[i]] += elvect[i]; }
      }
      for (int i = 0; i < F->ness; i++) { y[F->ess[i]] = 0.0; }
   }
   void orc_form_mult(const orc_form_t *F, const double *x, double *y)
   {
      orc_form_mult_range(F, x, y, 0, F->mesh.ne, 1);
   }

   // Sparsity: AddSubMatrix(vdofs, vdofs, elmat, skip_zeros=0) => full element
   // connectivity, explicit zeros kept.  Returned with sorted columns.
   // Call with rowptr==NULL to get nnz only.
   long orc_form_pattern(const orc_form_t *F, int *rowptr, int *colidx)
   {
      ElemCtx C(*F);
      const int N = C.goff.back();
      std::vector<std::vector<int>> rows(N);
      std::vector<int> vd;
      for (int e = 0; e < F->mesh.ne; e++)
      {
         element_vdofs(C, e, vd);
         for (int i = 0; i < C.nvd_total; i++)
         {
            std::vector<int> &r = rows[vd[i]];
            r.insert(r.end(), vd.begin(), vd.end());
         }
         if ((e & 1023) == 1023 || e == F->mesh.ne - 1)
         {
            // periodic compaction to bound memory
            for (int i = 0; i < C.nvd_total; i++)
            {
               std::vector<int> &r = rows[vd[i]];
               std::sort(r.begin(), r.end());
               r.erase(std::unique(r.begin(), r.end()), r.end());
            }
         }
      }
      long nnz = 0;
      for (int i = 0; i < N; i++)
      {
         std::vector<int> &r = rows[i];
         std::sort(r.begin(), r.end());
         r.erase(std::unique(r.begin(), r.end()), r.end());
         nnz += (long)r.size();
      }
      if (rowptr)
      {
         rowptr[0] = 0;
         for (int i = 0; i < N; i++)
         {
            rowptr[i + 1] = rowptr[i] + (int)rows[i].size();
            std::copy(rows[i].begin(), rows[i].end(), colidx + rowptr[i]);
         }
      }
      return nnz;
   }

   // NonlinearForm::GetGradient / BlockNonlinearForm::ComputeGradientBlocked
   // [MFEM-upstream]: AddSubMatrix in element order into the given (sorted) CSR
   // pattern, then EliminateRowCol(ess, DIAG_ONE).
   void orc_form_grad_range(const orc_form_t *F, const double *x, const int *rowptr, const int *colidx,
                            double *vals, int e0, int e1, int zero_vals)
   {
      ElemCtx C(*F);
      const int N = C.goff.back();
      if (zero_vals) { for (long i = 0; i < rowptr[N]; i++) { vals[i] = 0.0; } }
      std::vector<double> elmat((size_t)C.nvd_total * C.nvd_total);
      std::vector<int> vd;
      for (int e = e0; e < e1; e++)
      {
         element_grad(C, e, x, elmat.data());
         element_vdofs(C, e, vd);
         for (int i = 0; i < C.nvd_total; i++)
         {
            const int r = vd[i];
            const int *cb = colidx + rowptr[r], *ce = colidx + rowptr[r + 1];
            for (int j = 0; j < C.nvd_total; j++)
            {
               const int *p = std::lower_bound(cb, ce, vd[j]);
               vals[p - colidx] += elmat[i + (size_t)C.nvd_total * j];
            }
         }
      }
      for (int k = 0; k < F->ness; k++)
      {
         const int rc = F->ess[k];
         for (int p = rowptr[rc]; p < rowptr[rc + 1]; p++)
         {
            const int j = colidx[p];
            vals[p] = (j == rc) ? 1.0 : 0.0;
            if (j != rc)
            {
               const int *cb = colidx + rowptr[j], *ce = colidx + rowptr[j + 1];
               const int *pp = std::lower_bound(cb, ce, rc);
               if (pp != ce && *pp == rc) { vals[pp - colidx] = 0.0; }
            }
         }
      }
   }
   void orc_form_grad(const orc_form_t *F, const double *x, const int *rowptr, const int *colidx, double *vals)
   {
      orc_form_grad_range(F, x, rowptr, colidx, vals, 0, F->mesh.ne, 1);
   }

   // Interpolate per-point parameters / inputs (debug + DifferentiableCoefficient checks):
   // out[ne*nq*n_input] = x at each quadrature point
   void orc_form_inputs_at_qpts(const orc_form_t *F, const double *xg, double *out)
   {
      ElemCtx C(*F);
      std::vector<std::vector<double>> elfun, allshapes(F->nspaces);
      for (int s = 0; s < F->nspaces; s++) { allshapes[s].resize(C.dof[s] * C.sd[s]); }
      for (int e = 0; e < F->mesh.ne; e++)
      {
         gather(C, e, xg, elfun);
         for (int q = 0; q < C.ir.np; q++)
         {
            const double *ip = &C.ir.pts[q * C.dim];
            PointGeom g;
            eval_geom(C, e, ip, g);
            for (int s = 0; s < F->nspaces; s++) { calc_input_shapes(C, s, ip, q, g, allshapes[s].data()); }
            interp_inputs(C, allshapes, elfun, out + ((size_t)e * C.ir.np + q) * C.n_input);
         }
      }
   }

   // DifferentiableCoefficient::{Eval,Gradient,Hessian} projected to the rule's
   // points (src/ad_native.hpp:267-323; ex4.cpp:124-128,200): the form's spaces
   // supply the inputs (VALUE interpolation), fn is the wrapped ADFunction.
   // which: 0 value [ne*nq], 1 gradient [ne*nq*n], 2 hessian [ne*nq*n*n]
   void orc_form_coefficient(const orc_form_t *F, const double *xg, int which, double *out)
   {
      ElemCtx C(*F);
      const int n = C.n_input;
      std::vector<std::vector<double>> elfun, allshapes(F->nspaces);
      for (int s = 0; s < F->nspaces; s++) { allshapes[s].resize(C.dof[s] * C.sd[s]); }
      std::vector<double> x(n), qprm(std::max(C.nprm, 1));
      for (int e = 0; e < F->mesh.ne; e++)
      {
         gather(C, e, xg, elfun);
         for (int q = 0; q < C.ir.np; q++)
         {
            const double *ip = &C.ir.pts[q * C.dim];
            PointGeom g;
            eval_geom(C, e, ip, g);
            for (int s = 0; s < F->nspaces; s++) { calc_input_shapes(C, s, ip, q, g, allshapes[s].data()); }
            interp_inputs(C, allshapes, elfun, x.data());
            eval_params(C, e, q, ip, g, qprm.data());
            Ctx c {F->fn, qprm.data()};
            const size_t pt = (size_t)e * C.ir.np + q;
            if (which == 0) { out[pt] = fn_value(c, F->root, x.data()); }
            else if (which == 1) { fn_gradient(c, F->root, x.data(), out + pt * n); }
            else { fn_hessian(c, F->root, x.data(), out + pt * n * n); }
         }
      }
   }

   // ParametrizedFunctional::ParamGradient::Eval: src/mmto.cpp:4-38 (incl. SURVEY H6).
   // Form F: spaces[0] = the design GridFunction rho (VALUE, vdim = param_dim) used as the
   // parameter source of every f_i; params[0] = PRM_GF_GRAD of the state (VectorGradientGridFunction);
   // nodes[F->root] = the parent functional (reads its f_i values at qprm[n_state + i]);
   // fi[0..nfi) = node ids of the parameter functions f_i (ADFunctions of rho).
   // out[ne*nq*param_dim].
   void orc_mmto_param_gradient(const orc_form_t *F, const double *rho, int nfi, const int *fi, double *out)
   {
      ElemCtx C(*F);
      const int param_dim = C.n_input; // rho components
      const int n_state = C.F.params[0].size;
      std::vector<std::vector<double>> elfun, allshapes(F->nspaces);
      for (int s = 0; s < F->nspaces; s++) { allshapes[s].resize(C.dof[s] * C.sd[s]); }
      std::vector<double> rq(param_dim), val(n_state + nfi), dfdc(param_dim);
      for (int e = 0; e < F->mesh.ne; e++)
      {
         gather(C, e, rho, elfun);
         for (int q = 0; q < C.ir.np; q++)
         {
            const double *ip = &C.ir.pts[q * C.dim];
            PointGeom g;
            eval_geom(C, e, ip, g);
            for (int s = 0; s < F->nspaces; s++) { calc_input_shapes(C, s, ip, q, g, allshapes[s].data()); }
            interp_inputs(C, allshapes, elfun, rq.data()); // param_coeffs[i]'s evaluator: rho at the point
            double *J = out + ((size_t)e * C.ir.np + q) * param_dim;
            for (int j = 0; j < param_dim; j++) { J[j] = 0.0; }
            // parent.ProcessParameters: evaluate the f_i (:15); then the states (:17-20)
            Ctx cf {F->fn, nullptr};
            for (int i = 0; i < nfi; i++) { val[n_state + i] = fn_value(cf, fi[i], rq.data()); }
            eval_params(C, e, q, ip, g, val.data()); // state block at offset 0
            Ctx cp {F->fn, val.data()};
            for (int i = 0; i < nfi; i++)
            {
               fn_gradient(cf, fi[i], rq.data(), dfdc.data()); // param_coeffs[i]->Gradient().Eval (:27)
               const double keep = val[n_state + i];           // :29
               for (int j = 0; j < param_dim; j++)
               {
                  val[n_state + i] = dfdc[j];                  // :32
                  J[j] += fn_value(cp, F->root, val.data());   // :33  parent(state)
               }
               val[n_state + i] = keep;                        // :36
            }
         }
      }
   }

   // ADDofPGNonlinearFormIntegrator nodal terms: src/dof_pg.hpp:66-128 (vector), :131-231 (grad).
   // One primal/dual pair of scalar spaces with identical element dof maps (dof_pg.hpp:107-109).
   // The reference takes the weights from primal_fe.GetNodes() (zero in MFEM); here the
   // per-(element,node) weights w[e*nd + j] = Tr.Weight()*ip.weight are an explicit input, and the
   // entropy parameters are processed in both paths.
   // r_u, r_psi: global vectors (accumulated, element order); d_pp, d_up: diagonal Jacobian entries.
   void orc_dofpg_nodal(const orc_fn_t *nodes, int entropy, int ne, int nd, const int *e2l, const double *w,
                        double alpha, const double *u, const double *psi, const double *psik,
                        double *r_u, double *r_psi, double *d_pp, double *d_up)
   {
      Ctx c {nodes, nullptr};
      for (int e = 0; e < ne; e++)
      {
         for (int j = 0; j < nd; j++)
         {
            const int d = e2l[(size_t)e * nd + j];
            const double ww = w[(size_t)e * nd + j] / alpha; // :119
            double Jv, Hv;
            fn_gradient(c, entropy, &psi[d], &Jv);           // :123
            fn_hessian(c, entropy, &psi[d], &Hv);            // :223
            r_u[d] += (psi[d] - psik[d]) * ww;               // :124
            r_psi[d] += (u[d] - Jv) * ww;                    // :125 (assigned per element, then AddElementVector)
            d_pp[d] += -Hv * ww;                             // :226
            d_up[d] += ww;                                   // :227-228
         }
      }
   }
} // extern "C"
