// madb.hpp -- C++ host-side mirror of the reference's integrator / functional interface
// above the C ABI (include/mfemad_b200.h).  Same class names and argument meaning as the
// reference (src/_ad_intg.hpp, src/ad_native.hpp, src/pg.hpp, src/pg.cpp); MFEM is not
// available in this image, so Mesh / FiniteElementSpace / Vector are minimal stand-ins.
// Error behaviour follows the reference: MFEM_VERIFY / MFEM_ABORT print and abort
// (src/ad_native.hpp:167, src/pg.cpp:10-14).
#pragma once
#include "../../include/mfemad_b200.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#define MADB_VERIFY(cond, msg)                                                          \
   do {                                                                                 \
      if (!(cond)) { std::fprintf(stderr, "Verification failed: (%s) is false:\n --> %s\n", #cond, std::string(msg).c_str()); std::abort(); } \
   } while (0)
#define MADB_CALL(x) MADB_VERIFY((x) == 0, madb_last_error())

namespace madb_host
{
typedef double real_t;
typedef std::vector<double> Vector;

// ---- ADEval: src/_ad_intg.hpp:24-66 -------------------------------------------------
enum class ADEval
{
   QVALUE = 1 << 0, VALUE = 1 << 1, GRAD = 1 << 2, DIV = 1 << 3, CURL = 1 << 4,
   Hessian = 1 << 5, VECTOR = 1 << 6, VECFE = 1 << 7, NUMOPT = 1 << 8
};
constexpr ADEval operator|(ADEval a, ADEval b) { return static_cast<ADEval>(static_cast<int>(a) | static_cast<int>(b)); }
constexpr ADEval operator&(ADEval a, ADEval b) { return static_cast<ADEval>(static_cast<int>(a) & static_cast<int>(b)); }
constexpr ADEval operator~(ADEval m) { return static_cast<ADEval>(~static_cast<int>(m)); }
constexpr bool hasFlag(ADEval mode, ADEval flag) { return (mode & flag) == flag; }
template <ADEval mode> constexpr bool isValidADEval()
{
   if (static_cast<int>(mode & ADEval::Hessian) != 0) { return false; }
   if (hasFlag(mode, ADEval::QVALUE)) { return static_cast<int>(mode & (~(ADEval::QVALUE | ADEval::VECTOR))) == 0; }
   return true;
}

// ---- device context (one per rank) ----------------------------------------------------
class Device
{
public:
   madb_ctx *ctx = nullptr;
   explicit Device(int dev = 0) { MADB_CALL(madb_ctx_create(dev, &ctx)); }
   ~Device() { madb_ctx_destroy(ctx); }
   static Device &Get(int dev = 0) { static Device d(dev); return d; }
};

// ---- PGStepSizeRule: src/pg.hpp:10-34, src/pg.cpp:4-54 -----------------------------------
struct PGStepSizeRule
{
   enum RuleType { CONSTANT, POLY, EXP, DOUBLE_EXP, INVALID };
   RuleType rule_type;
   real_t max_alpha, alpha0, ratio, ratio2;
   PGStepSizeRule(int rule_type_, real_t alpha0_ = 1.0, real_t max_alpha_ = 1e06, real_t ratio_ = -1.0, real_t ratio2_ = -1.0)
      : rule_type(static_cast<RuleType>(rule_type_)), max_alpha(max_alpha_), alpha0(alpha0_), ratio(ratio_), ratio2(ratio2_)
   {
      MADB_VERIFY(rule_type_ < RuleType::INVALID, "PGStepSizeRule: Invalid rule type");
      MADB_VERIFY(alpha0 > 0, "PGStepSizeRule: alpha0 must be positive");
      MADB_VERIFY(max_alpha >= alpha0, "PGStepSizeRule: max_alpha must be greater than or equal to alpha0");
      if (rule_type == POLY) { MADB_VERIFY(ratio > 0, "PGStepSizeRule: ratio must be positive for POLY rule"); }
      else if (rule_type == EXP) { MADB_VERIFY(ratio > 1, "PGStepSizeRule: ratio must be greater than 1 for EXP rule"); }
      else if (rule_type == DOUBLE_EXP) { MADB_VERIFY(ratio > 1 && ratio2 > 1, "PGStepSizeRule: ratio and ratio2 must be greater than 1 for DOUBLE_EXP rule"); }
   }
   real_t Get(int iter) const
   {
      real_t alpha = alpha0;
      switch (rule_type)
      {
         case CONSTANT: break;
         case POLY: alpha *= std::pow(iter + 1, ratio); break;
         case EXP: alpha *= std::pow(ratio, iter); break;
         case DOUBLE_EXP: alpha *= std::pow(ratio, std::pow(ratio2, iter)); break;
         default: break;
      }
      return std::min(alpha, max_alpha);
   }
};

// ---- ADFunction handles: src/ad_native.hpp:137-190 -----------------------------------------
class ADFunction
{
protected:
   madb_functional *h = nullptr;
   std::vector<ADFunction *> children;
   void Create(const char *kind, const std::vector<double> &p, const std::vector<int> &ip = {})
   {
      std::vector<madb_functional *> ch;
      for (auto *c : children) { ch.push_back(c->Handle()); }
      MADB_CALL(madb_functional_create(Device::Get().ctx, kind, (int)p.size(), p.data(), (int)ip.size(), ip.data(),
                                       (int)ch.size(), ch.data(), &h));
   }
   virtual std::vector<double> Params() const { return {}; }
public:
   const int n_input;
   explicit ADFunction(int n_input_) : n_input(n_input_) {}
   virtual ~ADFunction() { madb_functional_destroy(h); }
   madb_functional *Handle() { return h; }
   /// push the current values of the (mutable) members to the device-side functional;
   /// the reference re-reads them at every quadrature point (e.g. ex2.cpp:98)
   virtual void Sync()
   {
      for (auto *c : children) { c->Sync(); }
      const std::vector<double> p = Params();
      MADB_CALL(madb_functional_set_params(h, (int)p.size(), p.data()));
   }
   /// value, gradient, Hessian at one point (src/ad_native.cpp:181-230), evaluated on the device
   real_t operator()(const Vector &x) { real_t v; Eval(x, &v, nullptr, nullptr); return v; }
   void Gradient(const Vector &x, Vector &J) { J.resize(x.size()); Eval(x, nullptr, J.data(), nullptr); }
   void Hessian(const Vector &x, Vector &H) { H.resize(x.size() * x.size()); Eval(x, nullptr, nullptr, H.data()); }
private:
   void Eval(const Vector &x, real_t *v, real_t *g, real_t *H)
   {
      MADB_VERIFY((int)x.size() == n_input, "ADFunction::operator(): var.Size() must match n_input");
      Sync();
      MADB_CALL(madb_functional_eval(Device::Get().ctx, h, n_input, 1, x.data(), nullptr, v, g, H));
   }
};

struct MassEnergy : ADFunction { explicit MassEnergy(int n_var) : ADFunction(n_var) { Create("mass", {}); } };
struct DiffusionEnergy : ADFunction { explicit DiffusionEnergy(int dim) : ADFunction(dim) { Create("diffusion", {}, {0}); } };
struct LinearElasticityEnergy : ADFunction
{
   real_t lambda, mu;
   LinearElasticityEnergy(int dim, real_t lambda_, real_t mu_) : ADFunction(dim * dim), lambda(lambda_), mu(mu_) { Create("elasticity", Params()); }
   std::vector<double> Params() const override { return {lambda, mu}; }
};
struct MinimalSurfaceEnergy : ADFunction // ex2.cpp:12-24
{
   real_t eps = 0.5;
   explicit MinimalSurfaceEnergy(int dim) : ADFunction(dim) { Create("minsurf", Params()); }
   std::vector<double> Params() const override { return {eps}; }
};
struct ObstacleEnergy : ADFunction { explicit ObstacleEnergy(int dim) : ADFunction(dim + 1) { Create("obstacle", {}); } };              // ex4.cpp:15-28
struct GradientObstacleEnergy : ADFunction { explicit GradientObstacleEnergy(int dim) : ADFunction(dim) { Create("gradobstacle", {}); } }; // ex5.cpp:15-22

// ---- entropies: src/pg.hpp:37-44, :253-376 ---------------------------------------------------
struct ADEntropy : ADFunction { using ADFunction::ADFunction; };
struct ShannonEntropy : ADEntropy
{
   real_t bound; int sign;
   ShannonEntropy(real_t bound_, int sign_ = 1) : ADEntropy(1), bound(bound_), sign(sign_)
   {
      MADB_VERIFY(sign == 1 || sign == -1, "ShannonEntropy: sign must be 1 or -1");
      Create("shannon", Params());
   }
   std::vector<double> Params() const override { return {bound, (double)sign}; }
};
struct FermiDiracEntropy : ADEntropy
{
   real_t lower, upper;
   FermiDiracEntropy(real_t lower_bound, real_t upper_bound) : ADEntropy(1), lower(lower_bound), upper(upper_bound) { Create("fermidirac", Params()); }
   std::vector<double> Params() const override { return {lower, upper}; }
};
struct HellingerEntropy : ADEntropy
{
   real_t bound;
   HellingerEntropy(int dim, real_t bound_) : ADEntropy(dim), bound(bound_) { Create("hellinger", Params()); }
   std::vector<double> Params() const override { return {bound}; }
};
struct SimplexEntropy : ADEntropy
{
   real_t bound;
   SimplexEntropy(int n, real_t bound_) : ADEntropy(n), bound(bound_) { Create("simplex", Params()); }
   std::vector<double> Params() const override { return {bound}; }
};

// ---- ADPGFunctional: src/pg.hpp:67-214 -------------------------------------------------------
class ADPGFunctional : public ADFunction
{
   ADFunction &f;
   ADEntropy &entropy;
   real_t alpha = 1.0;
public:
   ADPGFunctional(ADFunction &f_, ADEntropy &dual_entropy, int idx = 0)
      : ADFunction(f_.n_input + dual_entropy.n_input), f(f_), entropy(dual_entropy)
   {
      MADB_VERIFY(f.n_input >= idx + dual_entropy.n_input, "ADPGFunctional: f.n_input must not exceed primal_begin + dual_entropy.n_input");
      children = {&f, &entropy};
      Create("pg", Params(), {idx});
   }
   ADFunction &GetObjective() const { return f; }
   ADEntropy &GetEntropy() const { return entropy; }
   void SetAlpha(real_t a) { alpha = a; }
   real_t GetAlpha() const { return alpha; }
   std::vector<double> Params() const override { return {alpha}; }
};

// ---- minimal FE substrate ----------------------------------------------------------------------
class Mesh
{
public:
   int dim = 2, nx = 0, ny = 0;
   std::vector<int> e2n;
   std::vector<double> coords;
   madb_mesh *h = nullptr;
   static Mesh MakeCartesian2D(int nx, int ny, real_t sx = 1.0, real_t sy = 1.0)
   {
      Mesh m;
      m.nx = nx; m.ny = ny;
      for (int j = 0; j <= ny; j++) { for (int i = 0; i <= nx; i++) { m.coords.push_back(sx * i / nx); m.coords.push_back(sy * j / ny); } }
      for (int j = 0; j < ny; j++)
      {
         for (int i = 0; i < nx; i++)
         {
            const int v = j * (nx + 1) + i;
            for (int k : {v, v + 1, v + nx + 1, v + nx + 2}) { m.e2n.push_back(k); }
         }
      }
      MADB_CALL(madb_mesh_create(Device::Get().ctx, 2, nx * ny, m.e2n.data(), (nx + 1) * (ny + 1), m.coords.data(), &m.h));
      return m;
   }
   int GetNE() const { return nx * ny; }
   int Dimension() const { return dim; }
};

class FiniteElementSpace
{
public:
   Mesh &mesh;
   int basis, order, vdim, ndofs;
   std::vector<int> e2l;
   madb_space *h = nullptr;
   /// basis: MADB_BASIS_H1 (H1_FECollection) or MADB_BASIS_L2 (L2_FECollection)
   FiniteElementSpace(Mesh &m, int basis_, int order_, int vdim_ = 1) : mesh(m), basis(basis_), order(order_), vdim(vdim_)
   {
      const int p = order, n1 = p + 1;
      if (basis == MADB_BASIS_H1)
      {
         const int ngx = m.nx * p + 1, ngy = m.ny * p + 1;
         ndofs = ngx * ngy;
         for (int ey = 0; ey < m.ny; ey++) { for (int ex = 0; ex < m.nx; ex++) { for (int j = 0; j < n1; j++) { for (int i = 0; i < n1; i++) { e2l.push_back((ey * p + j) * ngx + ex * p + i); } } } }
      }
      else
      {
         ndofs = m.GetNE() * n1 * n1;
         for (int k = 0; k < ndofs; k++) { e2l.push_back(k); }
      }
      MADB_CALL(madb_space_create(Device::Get().ctx, m.h, basis, order, vdim, MADB_BYNODES, ndofs, e2l.data(), &h));
   }
   int GetVSize() const { return ndofs * vdim; }
};

struct SparseMatrix // CSR with sorted columns; stands where the drivers cast GetGradient to SparseMatrix& (ex1.cpp:64)
{
   std::vector<int> I, J;
   std::vector<double> A;
   int Height() const { return (int)I.size() - 1; }
};

// ---- integrators and forms -----------------------------------------------------------------
template <ADEval... modes> class ADBlockNonlinearFormIntegrator
{
public:
   static constexpr std::array<ADEval, sizeof...(modes)> modes_arr = {modes...};
   ADFunction &f;
   int quad_order;
   /// ir_order < 0: default rule 2*max_order+2 (src/_ad_intg.hpp:298-313); the reference passes an IntegrationRule*,
   /// here its ORDER (IntRules.Get(geom, order), ex4.cpp:104)
   explicit ADBlockNonlinearFormIntegrator(ADFunction &f_, int ir_order = -1) : f(f_), quad_order(ir_order)
   {
      static_assert((isValidADEval<modes>() && ...), "ADBlockNonlinearFormIntegrator: Invalid ADEval mode");
   }
};
template <ADEval mode> using ADNonlinearFormIntegrator = ADBlockNonlinearFormIntegrator<mode>;

class BlockNonlinearForm
{
   std::vector<FiniteElementSpace *> fes;
   std::vector<FiniteElementSpace *> pfes; // parameter GridFunction spaces (Evaluator sources)
   madb_integrator *intg = nullptr;
   ADFunction *fn = nullptr;
   SparseMatrix grad;
public:
   explicit BlockNonlinearForm(std::vector<FiniteElementSpace *> spaces) : fes(std::move(spaces)) {}
   ~BlockNonlinearForm()
   {
      if (solver) { madb_solver_destroy(solver); }
      if (d_vals) { madb_device_free(Device::Get().ctx, d_vals); }
      madb_integrator_destroy(intg);
   }
   /// GridFunction parameters of the functional's Evaluator (e.g. psi_k of ADPGFunctional, src/pg.hpp:106-111)
   void AddParameterSpace(FiniteElementSpace *s) { pfes.push_back(s); }
   template <ADEval... modes> void AddDomainIntegrator(ADBlockNonlinearFormIntegrator<modes...> *bfi)
   {
      MADB_VERIFY(sizeof...(modes) == fes.size(), "ADBlockNonlinearFormIntegrator: el.Size() must match numSpaces");
      std::vector<madb_space *> sp;
      std::vector<int> md, rl;
      for (size_t i = 0; i < fes.size(); i++) { sp.push_back(fes[i]->h); md.push_back(static_cast<int>(bfi->modes_arr[i])); rl.push_back(MADB_ROLE_INPUT); }
      for (auto *s : pfes) { sp.push_back(s->h); md.push_back(MADB_VALUE | (s->vdim > 1 ? MADB_VECTOR : 0)); rl.push_back(MADB_ROLE_PARAM); }
      fn = &bfi->f;
      MADB_CALL(madb_integrator_create(Device::Get().ctx, (int)sp.size(), sp.data(), md.data(), rl.data(), fn->Handle(), bfi->quad_order, &intg));
      delete bfi; // the form owns its integrators (ex1.cpp:55, ex4.cpp:139-142)
   }
   void SetParameter(int i, const Vector &gf) { MADB_CALL(madb_integrator_set_param_field(intg, (int)fes.size() + i, gf.data())); }
   void SetEssentialTrueDofs(const std::vector<int> &ess) { MADB_CALL(madb_integrator_set_essential(intg, (int)ess.size(), ess.data())); }
   int Height() const { int n = 0; for (auto *s : fes) { n += s->GetVSize(); } return n; }
   real_t GetEnergy(const Vector &x) { fn->Sync(); real_t e; MADB_CALL(madb_integrator_energy(intg, x.data(), &e)); return e; }
   void Mult(const Vector &x, Vector &y) { fn->Sync(); y.resize(x.size()); MADB_CALL(madb_integrator_mult(intg, x.data(), y.data())); }
   SparseMatrix &GetGradient(const Vector &x)
   {
      fn->Sync();
      if (grad.I.empty())
      {
         int64_t n, nnz;
         MADB_CALL(madb_integrator_pattern(intg, &n, &nnz, nullptr, nullptr));
         grad.I.resize(n + 1); grad.J.resize(nnz); grad.A.resize(nnz);
         MADB_CALL(madb_integrator_pattern(intg, nullptr, nullptr, grad.I.data(), grad.J.data()));
      }
      MADB_CALL(madb_integrator_grad_assemble(intg, x.data(), grad.A.data()));
      return grad;
   }
   /// Newton correction on the device: assembles the Jacobian at x into device memory and solves J(x) c = rhs with
   /// Jacobi-PCG (madb_solver_pcg; stands where the drivers hand the matrix to UMFPackSolver, ex1.cpp:64-66, ex2.cpp:80);
   /// the CSR values never reach the host.
   void SolveGradientPCG(const Vector &x, const Vector &rhs, Vector &c, real_t rtol = 1e-12, int maxit = 10000,
                         int *iters = nullptr, real_t *relres = nullptr)
   {
      fn->Sync();
      if (!solver) { MADB_CALL(madb_solver_create(intg, &solver)); }
      int64_t n = 0, nnz = 0;
      MADB_CALL(madb_integrator_pattern(intg, &n, &nnz, nullptr, nullptr));
      if (!d_vals) { MADB_VERIFY(madb_device_alloc(Device::Get().ctx, (size_t)nnz * sizeof(double), (void **)&d_vals) == 0, madb_last_error()); }
      MADB_CALL(madb_integrator_grad_assemble(intg, x.data(), d_vals));
      Vector sol(c.size() == (size_t)n ? c : Vector((size_t)n, 0.0));
      MADB_CALL(madb_solver_pcg(solver, d_vals, rhs.data(), sol.data(), rtol, 0.0, maxit, iters, relres));
      c = sol;
   }
private:
   madb_solver *solver = nullptr;
   double *d_vals = nullptr;
};
typedef BlockNonlinearForm NonlinearForm;

/// LinearForm + DomainLFIntegrator(f) of the drivers (ex4.cpp:145-148): b_i = (f, phi_i) on a scalar space, assembled on
/// the device as the residual of the energy f(x) u (functional "load"); f is sampled at the rule's points on the host
/// (Coefficient-type source).  ir_order < 0: MFEM's linear-form rule (order 2p).
class LinearForm : public Vector
{
   FiniteElementSpace &fes;
public:
   explicit LinearForm(FiniteElementSpace *s) : fes(*s) {}
   template <class F> void AddDomainIntegrator(F f, int ir_order = -1)
   {
      madb_functional *fn = nullptr;
      madb_integrator *intg = nullptr;
      madb_ctx *ctx = Device::Get().ctx;
      MADB_CALL(madb_functional_create(ctx, "load", 0, nullptr, 0, nullptr, 0, nullptr, &fn));
      madb_space *sp = fes.h;
      int md = MADB_VALUE, rl = MADB_ROLE_INPUT;
      const int ord = ir_order < 0 ? 2 * fes.order : ir_order, nq1 = (ord | 1) / 2 + 1; // IntRules.Get(geom, order)
      MADB_CALL(madb_integrator_create(ctx, 1, &sp, &md, &rl, fn, ord, &intg));
      const int64_t npts = (int64_t)fes.mesh.GetNE() * nq1 * nq1;
      std::vector<double> xyz((size_t)npts * fes.mesh.dim), qf((size_t)npts);
      MADB_CALL(madb_integrator_qpoint_coords(intg, xyz.data()));
      for (int64_t k = 0; k < npts; k++) { qf[k] = f(&xyz[(size_t)k * fes.mesh.dim]); }
      MADB_CALL(madb_integrator_set_param_qf(intg, 1, qf.data()));
      Vector zero((size_t)fes.GetVSize(), 0.0), b((size_t)fes.GetVSize(), 0.0);
      MADB_CALL(madb_integrator_mult(intg, zero.data(), b.data()));
      if (size() != b.size()) { assign(b.size(), 0.0); }
      for (size_t i = 0; i < b.size(); i++) { (*this)[i] += b[i]; }
      madb_integrator_destroy(intg);
      madb_functional_destroy(fn);
   }
   void Assemble() {} // the integrators assemble when they are added
};

} // namespace madb_host
