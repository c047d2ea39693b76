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
