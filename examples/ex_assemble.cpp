// Assembly drivers in the shape of the reference's examples, on the C++ host mirror:
//   ex0: AD known answers (ex0.cpp:100-163) on the device AD type
//   ex2: minimal-surface residual/Jacobian on a Cartesian mesh, eps halved between calls (ex2.cpp:94-99)
//   ex4: one PG block system H1(p+1) x L2(p-1) with FermiDirac entropy (ex4.cpp:99-142)
//   ex1: LinearForm + DomainLFIntegrator load vector (ex4.cpp:145-148) and one linear solve on the device
//        (diffusion Jacobian, Jacobi-PCG through madb_solver_pcg instead of UMFPackSolver, ex1.cpp:64-66)
// Prints checksums that tests/test_gpu_host_cpp.py compares with the Python path.
//   g++ -std=c++17 examples/ex_assemble.cpp -Lmfem-ad_b200 -lmadb -Wl,-rpath,$PWD/mfem-ad_b200 -o examples/ex_assemble
#include "../mfem-ad_b200/host/madb.hpp"
#include <cmath>
#include <cstdio>
using namespace madb_host;

static double dot(const Vector &a, const Vector &b) { double s = 0; for (size_t i = 0; i < a.size(); i++) { s += a[i] * b[i]; } return s; }

int main()
{
   {
      struct MyADFunction : ADFunction { MyADFunction() : ADFunction(3) { Create("ex0", {}); } } f;
      Vector x {0.5, 1.0, -1.0}, J, H;
      f.Gradient(x, J);
      f.Hessian(x, H);
      std::printf("ex0 value %.17g\n", f(x));
      std::printf("ex0 jac %.17g %.17g %.17g\n", J[0], J[1], J[2]);
      std::printf("ex0 hess %.17g %.17g %.17g %.17g\n", H[0], H[1], H[4], H[8]);
   }
   {
      Mesh mesh = Mesh::MakeCartesian2D(16, 12);
      FiniteElementSpace fes(mesh, MADB_BASIS_H1, 2);
      MinimalSurfaceEnergy energy(2);
      NonlinearForm nlf({&fes});
      nlf.AddDomainIntegrator(new ADNonlinearFormIntegrator<ADEval::GRAD>(energy));
      Vector x(fes.GetVSize()), y;
      for (size_t i = 0; i < x.size(); i++) { x[i] = std::sin(0.37 * i) * 0.3; }
      for (int it = 0; it < 2; it++)
      {
         nlf.Mult(x, y);
         SparseMatrix &K = nlf.GetGradient(x);
         double s = 0;
         for (double v : K.A) { s += v * v; }
         std::printf("ex2 eps=%.4f energy=%.17g y2=%.17g K2=%.17g nnz=%zu\n", energy.eps, nlf.GetEnergy(x), dot(y, y), s, K.A.size());
         energy.eps *= 0.5;
      }
   }
   {
      const int order = 2;
      Mesh mesh = Mesh::MakeCartesian2D(6, 5);
      FiniteElementSpace h1(mesh, MADB_BASIS_H1, order + 1), l2(mesh, MADB_BASIS_L2, order - 1);
      ObstacleEnergy obj(2);
      FermiDiracEntropy entropy(0.0, 0.5);
      ADPGFunctional pg(obj, entropy);
      BlockNonlinearForm bnlf({&h1, &l2});
      bnlf.AddParameterSpace(&l2);
      constexpr ADEval u_mode = ADEval::VALUE | ADEval::GRAD, psi_mode = ADEval::VALUE;
      bnlf.AddDomainIntegrator(new ADBlockNonlinearFormIntegrator<u_mode, psi_mode>(pg, 3 * order + 3));
      Vector x(bnlf.Height()), y, psik(l2.GetVSize());
      for (size_t i = 0; i < x.size(); i++) { x[i] = std::cos(0.11 * i) * 0.4; }
      for (size_t i = 0; i < psik.size(); i++) { psik[i] = std::sin(0.23 * i); }
      bnlf.SetParameter(0, psik);
      PGStepSizeRule rule(2, 0.1, 1e4, 2.0);
      pg.SetAlpha(rule.Get(3));
      bnlf.Mult(x, y);
      SparseMatrix &K = bnlf.GetGradient(x);
      double s = 0;
      for (double v : K.A) { s += v * v; }
      std::printf("ex4 alpha=%.4f y2=%.17g K2=%.17g nnz=%zu\n", pg.GetAlpha(), dot(y, y), s, K.A.size());
   }
   {
      Mesh mesh = Mesh::MakeCartesian2D(12, 12);
      FiniteElementSpace fes(mesh, MADB_BASIS_H1, 2);
      LinearForm b(&fes);
      b.AddDomainIntegrator([](const double *x) { return 2.0 * M_PI * M_PI * std::sin(M_PI * x[0]) * std::sin(M_PI * x[1]); });
      b.Assemble();
      double bs = 0;
      for (double v : b) { bs += v; }
      DiffusionEnergy energy(2);
      NonlinearForm nlf({&fes});
      nlf.AddDomainIntegrator(new ADNonlinearFormIntegrator<ADEval::GRAD>(energy));
      std::vector<int> ess;
      const int ng = 2 * 12 + 1;
      for (int j = 0; j < ng; j++) { for (int i = 0; i < ng; i++) { if (i == 0 || j == 0 || i == ng - 1 || j == ng - 1) { ess.push_back(j * ng + i); } } }
      nlf.SetEssentialTrueDofs(ess);
      for (int d : ess) { b[d] = 0.0; }
      Vector x(fes.GetVSize(), 0.0);
      int iters = 0;
      double relres = 0.0;
      nlf.SolveGradientPCG(x, b, x, 1e-12, 5000, &iters, &relres); // K(0) u = b on the device
      double err = 0.0;
      for (int j = 0; j < ng; j++)
      {
         for (int i = 0; i < ng; i++)
         {
            const double ex = std::sin(M_PI * i / (ng - 1.0)) * std::sin(M_PI * j / (ng - 1.0));
            err = std::max(err, std::fabs(x[j * ng + i] - ex));
         }
      }
      std::printf("ex1 bsum=%.17g iters=%d relres=%.3e maxerr=%.3e\n", bs, iters, relres, err);
   }
   return 0;
}
