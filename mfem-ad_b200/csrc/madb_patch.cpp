// madb_patch.cpp -- host-side setup of the patch assembly.
//
// The reference assembles with MFEM's element loop: AddElementVector and
// SparseMatrix::AddSubMatrix(vdofs, vdofs, elmat, skip_zeros=0) in element order
// (SURVEY a32).  On the GPU a per-element scatter writes every 8-byte value into
// its own 32-byte sector (profiles/r01_v1_k_element.md: 4x the algorithmic DRAM
// traffic).  Here elements are grouped into compact patches (recursive coordinate
// bisection, PATCH_PE elements = one CTA); a CTA accumulates complete CSR rows of
// its patch in shared memory and writes each row once, coalesced.
//   interior row  : every element containing the dof is in the patch -> final value
//   interface row : partial sums go to a staging buffer; a second kernel adds the
//                   partials of each entry in ascending patch order (deterministic).
// Inside a patch the elements are coloured (no two of a colour share a dof) and
// sorted by colour, so the shared-memory accumulation needs no atomics and has a
// fixed order.
#include "madb_host.hpp"

#include <algorithm>
#include <array>
#include <cstring>
#include <numeric>
#include <thread>

namespace madb
{

static int hw_threads_p()
{
   unsigned n = std::thread::hardware_concurrency();
   return (int)std::max(1u, std::min(n, 32u));
}
template <class Fn> static void parallel_for_p(long n, long serial_below, Fn fn)
{
   const int nt = (n < serial_below) ? 1 : hw_threads_p();
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

// recursive coordinate bisection into leaves of exactly pe elements (the last leaf may be smaller)
static void rcb(std::vector<int> &idx, int lo, int hi, const std::vector<double> &cen, int dim, int pe, int depth)
{
   const int n = hi - lo;
   if (n <= pe) { return; }
   double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
   for (int k = lo; k < hi; k++)
   {
      for (int d = 0; d < dim; d++)
      {
         const double c = cen[(size_t)idx[k] * 3 + d];
         mn[d] = std::min(mn[d], c);
         mx[d] = std::max(mx[d], c);
      }
   }
   int ax = 0;
   for (int d = 1; d < dim; d++) { if (mx[d] - mn[d] > mx[ax] - mn[ax]) { ax = d; } }
   const int npch = (n + pe - 1) / pe;
   const int left = (npch / 2) * pe;
   const int a1 = (ax + 1) % 3, a2 = (ax + 2) % 3;
   auto cmp = [&](int a, int b)
   {
      const double *ca = &cen[(size_t)a * 3], *cb = &cen[(size_t)b * 3];
      if (ca[ax] != cb[ax]) { return ca[ax] < cb[ax]; }
      if (ca[a1] != cb[a1]) { return ca[a1] < cb[a1]; }
      if (ca[a2] != cb[a2]) { return ca[a2] < cb[a2]; }
      return a < b;
   };
   std::nth_element(idx.begin() + lo, idx.begin() + lo + left, idx.begin() + hi, cmp);
   if (depth < 3)
   {
      std::thread t([&]() { rcb(idx, lo, lo + left, cen, dim, pe, depth + 1); });
      rcb(idx, lo + left, hi, cen, dim, pe, depth + 1);
      t.join();
   }
   else
   {
      rcb(idx, lo, lo + left, cen, dim, pe, depth + 1);
      rcb(idx, lo + left, hi, cen, dim, pe, depth + 1);
   }
}

// Patch order: fills I.perm (sorted position -> element), I.pdesc[p].ne/ncol/col_off.
void patch_order(Integrator &I)
{
   const int ne = I.ne, dim = I.mesh->dim, ngn = 1 << dim, pe = PATCH_PE;
   std::vector<double> cen((size_t)ne * 3, 0.0);
   for (int e = 0; e < ne; e++)
   {
      for (int k = 0; k < ngn; k++)
      {
         const int v = I.mesh->e2n[(size_t)e * ngn + k];
         for (int d = 0; d < dim; d++) { cen[(size_t)e * 3 + d] += I.mesh->coords[(size_t)v * dim + d]; }
      }
   }
   std::vector<int> idx(ne);
   std::iota(idx.begin(), idx.end(), 0);
   rcb(idx, 0, ne, cen, dim, pe, 0);
   const int np = (ne + pe - 1) / pe;
   I.pdesc.assign(np, PatchDesc());
   // colour and sort inside each patch
   bool ok = true;
   parallel_for_p(np, 64, [&](long b, long e)
   {
      std::vector<int> vd, dofs, col(pe), ord(pe);
      std::vector<unsigned> mask;
      std::vector<std::vector<int>> evd(pe);
      for (long p = b; p < e; p++)
      {
         const int lo = (int)p * pe, n = std::min(pe, ne - lo);
         // keep a reproducible order inside the patch before colouring
         std::sort(idx.begin() + lo, idx.begin() + lo + n);
         dofs.clear();
         for (int l = 0; l < n; l++)
         {
            build_vdofs(I, idx[lo + l], evd[l]);
            dofs.insert(dofs.end(), evd[l].begin(), evd[l].end());
         }
         std::sort(dofs.begin(), dofs.end());
         dofs.erase(std::unique(dofs.begin(), dofs.end()), dofs.end());
         mask.assign(dofs.size(), 0u);
         int ncol = 0;
         for (int l = 0; l < n; l++)
         {
            unsigned used = 0;
            for (int v : evd[l]) { used |= mask[std::lower_bound(dofs.begin(), dofs.end(), v) - dofs.begin()]; }
            int c = 0;
            while (c < 32 && (used >> c) & 1u) { c++; }
            if (c >= PATCH_MAXCOL) { ok = false; c = PATCH_MAXCOL - 1; }
            col[l] = c;
            ncol = std::max(ncol, c + 1);
            for (int v : evd[l]) { mask[std::lower_bound(dofs.begin(), dofs.end(), v) - dofs.begin()] |= 1u << c; }
         }
         PatchDesc &D = I.pdesc[p];
         std::memset(&D, 0, sizeof(D));
         D.ne = n;
         D.ncol = ncol;
         for (int l = 0; l < n; l++) { ord[l] = l; }
         std::stable_sort(ord.begin(), ord.begin() + n, [&](int a, int bb) { return col[a] < col[bb]; });
         std::vector<int> tmp(n);
         for (int l = 0; l < n; l++) { tmp[l] = idx[lo + ord[l]]; }
         int k = 0;
         for (int c = 0; c <= PATCH_MAXCOL; c++)
         {
            while (k < n && col[ord[k]] < c) { k++; }
            D.col_off[c] = (unsigned char)k;
         }
         for (int c = ncol; c <= PATCH_MAXCOL; c++) { D.col_off[c] = (unsigned char)n; }
         std::copy(tmp.begin(), tmp.end(), idx.begin() + lo);
      }
   });
   if (!ok) { I.pdesc.clear(); I.use_patches = false; return; }
   I.perm = idx;
}

// Residual side: local rows per patch, interior/interface split, yslot map, interface reduction lists.
void patch_build_y(Integrator &I, PatchHostY &H)
{
   const int ne = I.ne, nvd = I.nvd, pe = PATCH_PE, np = (int)I.pdesc.size();
   // which patches touch a dof: first patch id, or -2 when more than one
   std::vector<int> owner(I.ntotal, -1);
   {
      std::vector<int> vd;
      for (int t = 0; t < ne; t++)
      {
         const int p = t / pe;
         build_vdofs(I, I.perm[t], vd);
         for (int v : vd)
         {
            if (owner[v] == -1) { owner[v] = p; }
            else if (owner[v] != p) { owner[v] = -2; }
         }
      }
   }
   I.prow_off.assign(np + 1, 0);
   std::vector<std::vector<int>> rows(np);
   H.yslot.assign((size_t)nvd * I.stride, 0);
   parallel_for_p(np, 64, [&](long b, long e)
   {
      std::vector<int> vd, dofs, ifc;
      for (long p = b; p < e; p++)
      {
         PatchDesc &D = I.pdesc[p];
         const int lo = (int)p * pe;
         dofs.clear();
         for (int l = 0; l < D.ne; l++)
         {
            build_vdofs(I, I.perm[lo + l], vd);
            dofs.insert(dofs.end(), vd.begin(), vd.end());
         }
         std::sort(dofs.begin(), dofs.end());
         dofs.erase(std::unique(dofs.begin(), dofs.end()), dofs.end());
         std::vector<int> &R = rows[p];
         R.clear();
         ifc.clear();
         for (int v : dofs) { (owner[v] == -2 ? ifc : R).push_back(v); }
         D.nrow_int = (int)R.size();
         R.insert(R.end(), ifc.begin(), ifc.end());
         D.nrows = (int)R.size();
         for (int l = 0; l < D.ne; l++)
         {
            build_vdofs(I, I.perm[lo + l], vd);
            for (int i = 0; i < nvd; i++)
            {
               const int v = vd[i];
               int lr;
               if (owner[v] != -2) { lr = (int)(std::lower_bound(R.begin(), R.begin() + D.nrow_int, v) - R.begin()); }
               else { lr = (int)(std::lower_bound(R.begin() + D.nrow_int, R.end(), v) - R.begin()); }
               H.yslot[(size_t)i * I.stride + lo + l] = (unsigned short)patch_swz(lr);
            }
         }
      }
   });
   // offsets, lists
   I.max_rows = 0;
   long yoff = 0, soff = 0;
   for (int p = 0; p < np; p++)
   {
      PatchDesc &D = I.pdesc[p];
      D.y_off = (int)yoff;
      D.ystage_off = (int)soff;
      yoff += D.nrow_int;
      soff += D.nrows - D.nrow_int;
      I.prow_off[p + 1] = I.prow_off[p] + D.nrows;
      I.max_rows = std::max(I.max_rows, D.nrows);
   }
   H.ystage_size = soff;
   H.ylist.resize(std::max<long>(yoff, 1));
   I.prows.resize(I.prow_off[np]);
   std::vector<std::pair<int, int>> tup; // (dof, stage index) in ascending patch order
   tup.reserve(soff);
   for (int p = 0; p < np; p++)
   {
      const PatchDesc &D = I.pdesc[p];
      std::copy(rows[p].begin(), rows[p].end(), I.prows.begin() + I.prow_off[p]);
      std::copy(rows[p].begin(), rows[p].begin() + D.nrow_int, H.ylist.begin() + D.y_off);
      for (int k = D.nrow_int; k < D.nrows; k++) { tup.emplace_back(rows[p][k], D.ystage_off + (k - D.nrow_int)); }
   }
   std::stable_sort(tup.begin(), tup.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.first < b.first; });
   H.y_ptr.assign(1, 0);
   H.y_dst.clear();
   H.y_src.resize(tup.size());
   for (size_t k = 0; k < tup.size(); k++)
   {
      if (k == 0 || tup[k].first != tup[k - 1].first)
      {
         if (k) { H.y_ptr.push_back((int)k); }
         H.y_dst.push_back(tup[k].first);
      }
      H.y_src[k] = tup[k].second;
   }
   H.y_ptr.push_back((int)tup.size());
   if (tup.empty()) { H.y_ptr.assign(1, 0); }
}

// Matrix side: slot of every element-matrix entry, runs of interior slots, interface reduction lists.
void patch_build_v(Integrator &I, PatchHostV &H)
{
   const int nvd = I.nvd, pe = PATCH_PE, np = (int)I.pdesc.size();
   H.pslot.assign((size_t)nvd * nvd * I.stride, 0);
   struct PerPatch
   {
      std::vector<int> run_s, run_g;
      std::vector<int> ifc_gpos; // global CSR position of each interface slot
   };
   std::vector<PerPatch> PP(np);
   bool ok = true;
   parallel_for_p(np, 64, [&](long b, long e)
   {
      std::vector<int> vd, base;
      std::vector<long> keys;
      for (long p = b; p < e; p++)
      {
         PatchDesc &D = I.pdesc[p];
         PerPatch &P = PP[p];
         const int lo = (int)p * pe;
         const int *R = I.prows.data() + I.prow_off[p];
         // interior rows: slots follow the CSR rows; merge rows with consecutive dof ids into runs
         base.assign(D.nrows, 0);
         int s = 0;
         P.run_s.clear(); P.run_g.clear();
         for (int lr = 0; lr < D.nrow_int; lr++)
         {
            const int r = R[lr];
            if (lr == 0 || R[lr - 1] + 1 != r) { P.run_s.push_back(s); P.run_g.push_back(I.rowptr[r]); }
            base[lr] = s;
            s += I.rowptr[r + 1] - I.rowptr[r];
         }
         D.nint = s;
         D.nruns = (int)P.run_s.size();
         P.run_s.push_back(s); // sentinel
         P.run_g.push_back(0);
         // interface rows: the columns present in this patch
         keys.clear();
         for (int l = 0; l < D.ne; l++)
         {
            build_vdofs(I, I.perm[lo + l], vd);
            for (int i = 0; i < nvd; i++)
            {
               const int lr = (int)(std::lower_bound(R + D.nrow_int, R + D.nrows, vd[i]) - R);
               if (lr < D.nrows && R[lr] == vd[i] && lr >= D.nrow_int)
               {
                  for (int j = 0; j < nvd; j++) { keys.push_back(((long)lr << 32) | (unsigned)vd[j]); }
               }
            }
         }
         std::sort(keys.begin(), keys.end());
         keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
         D.nslots = D.nint + (int)keys.size();
         if (D.nslots > 65520) { ok = false; }
         P.ifc_gpos.resize(keys.size());
         for (size_t k = 0; k < keys.size(); k++)
         {
            const int r = R[(int)(keys[k] >> 32)], c = (int)(keys[k] & 0xffffffff);
            const int *cb = I.colidx.data() + I.rowptr[r], *ce = I.colidx.data() + I.rowptr[r + 1];
            P.ifc_gpos[k] = (int)(std::lower_bound(cb, ce, c) - I.colidx.data());
         }
         // element entries -> slots
         for (int l = 0; l < D.ne; l++)
         {
            build_vdofs(I, I.perm[lo + l], vd);
            for (int i = 0; i < nvd; i++)
            {
               int lr = (int)(std::lower_bound(R, R + D.nrow_int, vd[i]) - R);
               const bool interior = lr < D.nrow_int && R[lr] == vd[i];
               if (!interior) { lr = (int)(std::lower_bound(R + D.nrow_int, R + D.nrows, vd[i]) - R); }
               const int r = vd[i];
               const int *cb = I.colidx.data() + I.rowptr[r], *ce = I.colidx.data() + I.rowptr[r + 1];
               for (int j = 0; j < nvd; j++)
               {
                  int slot;
                  if (interior) { slot = base[lr] + (int)(std::lower_bound(cb, ce, vd[j]) - cb); }
                  else
                  {
                     const long key = ((long)lr << 32) | (unsigned)vd[j];
                     slot = D.nint + (int)(std::lower_bound(keys.begin(), keys.end(), key) - keys.begin());
                  }
                  H.pslot[((size_t)i * nvd + j) * I.stride + lo + l] = (unsigned short)patch_swz(slot);
               }
            }
         }
      }
   });
   if (!ok) { set_error("patch assembly: more than 65535 matrix slots in one patch"); I.have_patch_vals = false; return; }
   long roff = 0, soff = 0;
   I.max_slots = 0;
   I.max_runs = 0;
   for (int p = 0; p < np; p++)
   {
      PatchDesc &D = I.pdesc[p];
      D.run_off = (int)roff;
      D.stage_off = (int)soff;
      roff += D.nruns + 1;
      soff += D.nslots - D.nint;
      I.max_slots = std::max(I.max_slots, D.nslots);
      I.max_runs = std::max(I.max_runs, D.nruns);
   }
   H.vstage_size = soff;
   H.run_s.resize(roff); H.run_g.resize(roff); 
   std::vector<std::pair<int, int>> tup((size_t)soff);
   parallel_for_p(np, 64, [&](long b, long e)
   {
      for (long p = b; p < e; p++)
      {
         const PatchDesc &D = I.pdesc[p];
         std::copy(PP[p].run_s.begin(), PP[p].run_s.end(), H.run_s.begin() + D.run_off);
         std::copy(PP[p].run_g.begin(), PP[p].run_g.end(), H.run_g.begin() + D.run_off);
         for (size_t k = 0; k < PP[p].ifc_gpos.size(); k++) { tup[(size_t)D.stage_off + k] = {PP[p].ifc_gpos[k], D.stage_off + (int)k}; }
      }
   });
   // group by CSR position, sources in ascending staging (= patch) order
   std::stable_sort(tup.begin(), tup.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.first < b.first; });
   H.v_ptr.clear(); H.v_dst.clear();
   H.v_src.resize(tup.size());
   for (size_t k = 0; k < tup.size(); k++)
   {
      if (k == 0 || tup[k].first != tup[k - 1].first)
      {
         H.v_ptr.push_back((int)k);
         H.v_dst.push_back(tup[k].first);
      }
      H.v_src[k] = tup[k].second;
   }
   H.v_ptr.push_back((int)tup.size());
   I.have_patch_vals = true;
}

} // namespace madb
