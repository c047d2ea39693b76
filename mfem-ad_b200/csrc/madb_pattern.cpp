// madb_pattern.cpp -- host-side setup: 1-D bases/rules, element colouring,
// CSR sparsity of the full element connectivity (MFEM AddSubMatrix(...,
// skip_zeros=0)) and the element -> CSR scatter map.
#include "madb_host.hpp"

#include <algorithm>
#include <cmath>
#include <thread>

namespace madb
{

// ---------------------------------------------------------------------------
// Gauss-Legendre / Gauss-Lobatto points on [0,1] by Newton iteration on the
// Legendre three-term recurrence, and barycentric Lagrange tables.
// ---------------------------------------------------------------------------
static void legendre(int n, double z, double &p, double &dp)
{
   // P_n(z) and P_n'(z)
   double pm1 = 1.0, pm0 = z;
   if (n == 0) { p = 1.0; dp = 0.0; return; }
   for (int k = 2; k <= n; k++)
   {
      const double pk = ((2 * k - 1) * z * pm0 - (k - 1) * pm1) / k;
      pm1 = pm0;
      pm0 = pk;
   }
   p = pm0;
   dp = n * (z * pm0 - pm1) / (z * z - 1.0);
}

void gauss_legendre_01(int n, std::vector<double> &x, std::vector<double> &w)
{
   x.assign(n, 0.0);
   w.assign(n, 0.0);
   for (int i = 0; i < (n + 1) / 2; i++)
   {
      double z = std::cos(M_PI * (i + 0.75) / (n + 0.5)), p = 0, dp = 1;
      for (int it = 0; it < 64; it++)
      {
         legendre(n, z, p, dp);
         const double dz = p / dp;
         z -= dz;
         if (std::fabs(dz) < 2e-16) { break; }
      }
      legendre(n, z, p, dp);
      const double wi = 1.0 / ((1.0 - z * z) * dp * dp); // weight on [0,1] = half the [-1,1] weight
      x[i] = 0.5 * (1.0 - z);
      x[n - 1 - i] = 0.5 * (1.0 + z);
      w[i] = w[n - 1 - i] = wi;
   }
   if (n % 2 == 1) { x[n / 2] = 0.5; }
}

void gauss_lobatto_01(int n, std::vector<double> &x)
{
   x.assign(n, 0.0);
   if (n == 1) { x[0] = 0.5; return; }
   x[n - 1] = 1.0;
   const int N = n - 1; // interior nodes: zeros of P_N'
   for (int i = 1; i <= (n - 2 + 1) / 2; i++)
   {
      double z = std::cos(M_PI * i / N); // descending from 1
      for (int it = 0; it < 64; it++)
      {
         double p, dp;
         legendre(N, z, p, dp);
         const double ddp = (2.0 * z * dp - N * (N + 1.0) * p) / (1.0 - z * z);
         const double dz = dp / ddp;
         z -= dz;
         if (std::fabs(dz) < 2e-16) { break; }
      }
      x[n - 1 - i] = 0.5 * (1.0 + z);
      x[i] = 0.5 * (1.0 - z);
   }
   if (n % 2 == 1) { x[n / 2] = 0.5; }
}

void lagrange_tables(const std::vector<double> &nodes, const std::vector<double> &pts,
                     std::vector<double> &B, std::vector<double> &G)
{
   const int nn = (int)nodes.size(), np = (int)pts.size();
   std::vector<double> bw(nn, 1.0); // barycentric weights
   for (int j = 0; j < nn; j++)
   {
      for (int k = 0; k < nn; k++) { if (k != j) { bw[j] /= (nodes[j] - nodes[k]); } }
   }
   B.assign((size_t)np * nn, 0.0);
   G.assign((size_t)np * nn, 0.0);
   for (int q = 0; q < np; q++)
   {
      const double t = pts[q];
      for (int j = 0; j < nn; j++)
      {
         double val = bw[j], der = 0.0;
         for (int k = 0; k < nn; k++) { if (k != j) { val *= (t - nodes[k]); } }
         for (int m = 0; m < nn; m++)
         {
            if (m == j) { continue; }
            double pr = bw[j];
            for (int k = 0; k < nn; k++) { if (k != j && k != m) { pr *= (t - nodes[k]); } }
            der += pr;
         }
         B[(size_t)q * nn + j] = val;
         G[(size_t)q * nn + j] = der;
      }
   }
}

// ---------------------------------------------------------------------------
// element vdofs in the concatenated numbering, element-vector order
// [field][component][dof]  (input fields only)
// ---------------------------------------------------------------------------
void build_vdofs(const Integrator &I, int e, std::vector<int> &vd)
{
   vd.resize(I.nvd);
   int k = 0, blk = 0;
   for (const FieldDesc &f : I.fields)
   {
      if (f.role != 0) { continue; }
      const Space &S = *f.space;
      const int nd = S.nd_el();
      const long off = I.goff[blk++];
      for (int c = 0; c < S.vdim; c++)
      {
         for (int i = 0; i < nd; i++)
         {
            const int d = S.e2l[(size_t)e * nd + i];
            vd[k++] = (int)(off + (S.ordering == ORD_BYNODES ? (long)d + (long)S.ndofs * c : (long)d * S.vdim + c));
         }
      }
   }
}

// dof -> elements adjacency over all input vdofs
static void build_dof2elem(const Integrator &I, std::vector<int> &ptr, std::vector<int> &lst)
{
   const long N = I.ntotal;
   ptr.assign(N + 1, 0);
   std::vector<int> vd;
   for (int e = 0; e < I.ne; e++)
   {
      build_vdofs(I, e, vd);
      for (int v : vd) { ptr[v + 1]++; }
   }
   for (long i = 0; i < N; i++) { ptr[i + 1] += ptr[i]; }
   lst.resize(ptr[N]);
   std::vector<int> fill(ptr.begin(), ptr.end() - 1);
   for (int e = 0; e < I.ne; e++)
   {
      build_vdofs(I, e, vd);
      for (int v : vd) { lst[fill[v]++] = e; }
   }
}

// greedy colouring: no two elements of a colour share a dof
void color_elements(const Integrator &I, std::vector<int> &color, int &ncolors)
{
   std::vector<int> ptr, lst;
   build_dof2elem(I, ptr, lst);
   color.assign(I.ne, -1);
   ncolors = 0;
   std::vector<int> vd;
   std::vector<char> used;
   for (int e = 0; e < I.ne; e++)
   {
      build_vdofs(I, e, vd);
      used.assign(ncolors + 1, 0);
      for (int v : vd)
      {
         for (int p = ptr[v]; p < ptr[v + 1]; p++)
         {
            const int c = color[lst[p]];
            if (c >= 0) { used[c] = 1; }
         }
      }
      int c = 0;
      while (c < ncolors && used[c]) { c++; }
      color[e] = c;
      if (c == ncolors) { ncolors++; }
   }
}

static int hw_threads()
{
   unsigned n = std::thread::hardware_concurrency();
   return (int)std::max(1u, std::min(n, 32u));
}

template <class Fn> static void parallel_for(long n, Fn fn)
{
   const int nt = (n < 4096) ? 1 : hw_threads();
   if (nt == 1) { fn(0, n); return; }
   std::vector<std::thread> th;
   const long chunk = (n + nt - 1) / nt;
   for (int t = 0; t < nt; t++)
   {
      const long b = t * chunk, e = std::min(n, b + chunk);
      if (b < e) { th.emplace_back([=]() { fn(b, e); }); }
   }
   for (auto &t : th) { t.join(); }
}

// CSR pattern with sorted columns: row v couples to every vdof of every element containing v
void build_pattern(Integrator &I)
{
   if (I.have_pattern) { return; }
   std::vector<int> ptr, lst;
   build_dof2elem(I, ptr, lst);
   const long N = I.ntotal;
   I.rowptr.assign(N + 1, 0);
   // pass 1: row lengths
   std::vector<int> len(N, 0);
   parallel_for(N, [&](long b, long e)
   {
      std::vector<int> cols, vd;
      for (long r = b; r < e; r++)
      {
         cols.clear();
         for (int p = ptr[r]; p < ptr[r + 1]; p++)
         {
            build_vdofs(I, lst[p], vd);
            cols.insert(cols.end(), vd.begin(), vd.end());
         }
         std::sort(cols.begin(), cols.end());
         len[r] = (int)(std::unique(cols.begin(), cols.end()) - cols.begin());
      }
   });
   long nnz = 0;
   for (long r = 0; r < N; r++)
   {
      nnz += len[r];
      if (nnz >= 0x7fffffffL) { set_error("Jacobian has >= 2^31 nonzeros: use the matrix-free action (grad_mult)"); I.rowptr.clear(); return; }
      I.rowptr[r + 1] = (int)nnz;
   }
   I.colidx.resize(nnz);
   parallel_for(N, [&](long b, long e)
   {
      std::vector<int> cols, vd;
      for (long r = b; r < e; r++)
      {
         cols.clear();
         for (int p = ptr[r]; p < ptr[r + 1]; p++)
         {
            build_vdofs(I, lst[p], vd);
            cols.insert(cols.end(), vd.begin(), vd.end());
         }
         std::sort(cols.begin(), cols.end());
         cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
         std::copy(cols.begin(), cols.end(), I.colidx.begin() + I.rowptr[r]);
      }
   });
   I.have_pattern = true;
}

// e2csr[(i*nvd+j)*stride + t] = position of (vd[i], vd[j]) of sorted element t,
// bit 31 set on the first (lowest-colour) contribution to that position.
void build_e2csr(const Integrator &I, const std::vector<int> &color, std::vector<int> &e2csr)
{
   const int nvd = I.nvd;
   e2csr.assign((size_t)nvd * nvd * I.stride, 0);
   parallel_for(I.ne, [&](long b, long e)
   {
      std::vector<int> vd;
      for (long t = b; t < e; t++)
      {
         build_vdofs(I, I.perm[t], vd);
         for (int i = 0; i < nvd; i++)
         {
            const int r = vd[i];
            const int *cb = I.colidx.data() + I.rowptr[r], *ce = I.colidx.data() + I.rowptr[r + 1];
            for (int j = 0; j < nvd; j++)
            {
               const int *p = std::lower_bound(cb, ce, vd[j]);
               e2csr[((size_t)i * nvd + j) * I.stride + t] = (int)(p - I.colidx.data());
            }
         }
      }
   });
   // first-touch flags: colours ascending; inside one colour no two elements share an entry
   std::vector<unsigned char> touched(I.colidx.size(), 0);
   const int ncolors = (int)I.color_off.size() - 1;
   for (int c = 0; c < ncolors; c++)
   {
      parallel_for(I.color_off[c + 1] - I.color_off[c], [&](long b, long e)
      {
         for (long t = I.color_off[c] + b; t < I.color_off[c] + e; t++)
         {
            for (int ij = 0; ij < nvd * nvd; ij++)
            {
               int &m = e2csr[(size_t)ij * I.stride + t];
               if (!touched[m]) { touched[m] = 1; m |= 0x80000000; }
            }
         }
      });
   }
   (void)color;
}

} // namespace madb

// ---------------------------------------------------------------------------------------------
// Triangles: MFEM's integration rules and nodal bases (restated from the published tables: Strang-Fix / Dunavant points
// as MFEM's IntegrationRules::TriangleIntegrationRule lists them, point order of AddTriMidPoint / AddTriPoints3 /
// AddTriPoints6).  tests/test_simplex.py checks the polynomial exactness of every rule, which pins the digits.
// ---------------------------------------------------------------------------------------------
namespace madb
{
namespace
{
void tri_mid(std::vector<double> &p, std::vector<double> &w, double weight)
{
   p.push_back(1.0 / 3.0); p.push_back(1.0 / 3.0);
   w.push_back(weight);
}
void tri3(std::vector<double> &p, std::vector<double> &w, double a, double weight)
{
   const double b = 1.0 - 2.0 * a;
   const double xy[3][2] = {{a, a}, {a, b}, {b, a}};
   for (auto &q : xy) { p.push_back(q[0]); p.push_back(q[1]); w.push_back(weight); }
}
void tri6(std::vector<double> &p, std::vector<double> &w, double a, double b, double weight)
{
   const double c = 1.0 - a - b;
   const double xy[6][2] = {{a, b}, {b, a}, {a, c}, {c, a}, {b, c}, {c, b}};
   for (auto &q : xy) { p.push_back(q[0]); p.push_back(q[1]); w.push_back(weight); }
}
} // namespace

bool triangle_rule(int order, std::vector<double> &p, std::vector<double> &w)
{
   p.clear();
   w.clear();
   switch (order)
   {
      case 0:
      case 1: tri_mid(p, w, 0.5); break;
      case 2: tri3(p, w, 1.0 / 6.0, 1.0 / 6.0); break;
      case 3:
         tri_mid(p, w, -0.28125);
         tri3(p, w, 0.2, 25.0 / 96.0);
         break;
      case 4:
         tri3(p, w, 0.091576213509770743460, 0.054975871827660933819);
         tri3(p, w, 0.44594849091596488632, 0.11169079483900573285);
         break;
      case 5:
         tri_mid(p, w, 0.1125);
         tri3(p, w, 0.10128650732345633880, 0.062969590272413576298);
         tri3(p, w, 0.47014206410511508977, 0.066197076394253090369);
         break;
      case 6:
         tri3(p, w, 0.063089014491502228340, 0.025422453185103408460);
         tri3(p, w, 0.24928674517091042129, 0.058393137863189683013);
         tri6(p, w, 0.053145049844816947353, 0.31035245103378440542, 0.041425537809186787597);
         break;
      default: return false;
   }
   return true;
}

int triangle_ndof(int basis, int order)
{
   if (basis == BASIS_H1) { return order == 1 ? 3 : (order == 2 ? 6 : 0); }
   return order == 0 ? 1 : 0;
}

void triangle_shapes(int basis, int order, double x, double y, double *phi, double *dphi)
{
   if (basis != BASIS_H1)
   {
      phi[0] = 1.0;
      dphi[0] = dphi[1] = 0.0;
      return;
   }
   const double l[3] = {1.0 - x - y, x, y};
   const double dl[3][2] = {{-1.0, -1.0}, {1.0, 0.0}, {0.0, 1.0}};
   if (order == 1)
   {
      for (int i = 0; i < 3; i++) { phi[i] = l[i]; dphi[2 * i] = dl[i][0]; dphi[2 * i + 1] = dl[i][1]; }
      return;
   }
   for (int i = 0; i < 3; i++)
   {
      phi[i] = l[i] * (2.0 * l[i] - 1.0);
      for (int k = 0; k < 2; k++) { dphi[2 * i + k] = (4.0 * l[i] - 1.0) * dl[i][k]; }
   }
   const int ed[3][2] = {{0, 1}, {1, 2}, {2, 0}};
   for (int e = 0; e < 3; e++)
   {
      const int a = ed[e][0], b = ed[e][1];
      phi[3 + e] = 4.0 * l[a] * l[b];
      for (int k = 0; k < 2; k++) { dphi[2 * (3 + e) + k] = 4.0 * (dl[a][k] * l[b] + l[a] * dl[b][k]); }
   }
}
} // namespace madb
