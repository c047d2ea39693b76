// madb_patch.cpp -- host-side setup of the patch assembly.
//
// The reference assembles with MFEM's element loop: AddElementVector and
// SparseMatrix::AddSubMatrix(vdofs, vdofs, elmat, skip_zeros=0) in element order
// (SURVEY a32).  On the GPU a per-element scatter writes every 8-byte value into
// its own 32-byte sector (profiles/r01_v1_k_element.md: 4x the algorithmic DRAM
// traffic).  Here elements are grouped into compact patches (recursive coordinate
// bisection, PATCH_PE elements = one CTA).  The CTA stages its element matrices in
// shared memory; every CSR entry ("slot") of the patch is then gathered from its
// sources in ascending element order and written once, coalesced.
//   interior row  : every element containing the dof is in the patch -> final value
//   interface row : partial sums go to a staging buffer; a second kernel adds the
//                   partials of each entry in ascending patch order (deterministic).
// No atomics, no colouring: the summation order of every entry is fixed by the maps.
#include "madb_host.hpp"
#include "madb_pair_schedule.hpp"

#include <algorithm>
#include <cstdio>
#include <array>
#include <cmath>
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

// Patch order: fills I.perm (sorted position -> element) and I.pdesc[p].ne.
void patch_order(Integrator &I)
{
   const int ne = I.ne, dim = I.mesh->dim, ngn = I.mesh->ngn(), pe = I.pe;
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
   for (int p = 0; p < np; p++)
   {
      std::memset(&I.pdesc[p], 0, sizeof(PatchDesc));
      const int lo = p * pe, n = std::min(pe, ne - lo);
      I.pdesc[p].ne = n;
      std::sort(idx.begin() + lo, idx.begin() + lo + n); // ascending element id = summation order inside the patch
   }
   I.perm = idx;
}

namespace
{
struct BlobWriter
{
   std::vector<unsigned char> b;
   template <class T> void section(const std::vector<T> &v, size_t n)
   {
      const size_t bytes = n * sizeof(T), off = b.size();
      b.resize(off + (size_t)patch_al16((int)bytes), 0);
      if (bytes) { std::memcpy(b.data() + off, v.data(), bytes); }
   }
};

// Sources per destination (shared-memory locations, ascending element order) -> first source per destination
// plus the fold list: in phase k (k >= 1) the k-th source of every destination is added onto its first source,
// so afterwards the first source holds the complete sum (ascending element order for every destination).  Destinations with identical source lists (the (i,j)
// and (j,i) entries of a symmetric element matrix) share their folds.  Layout of `fold`:
// 8 counts (phases 1..8; u32) followed by the (dst | src << 16) words, phase by phase.
bool pack_sources(const std::vector<std::vector<unsigned short>> &srcs, int ndst, std::vector<unsigned short> &first,
                  std::vector<unsigned> &fold)
{
   first.assign(ndst, 0);
   // one record per distinct first source: (number of further sources, first source, list)
   std::vector<std::pair<std::pair<int, unsigned short>, const std::vector<unsigned short> *>> rec;
   for (int d = 0; d < ndst; d++)
   {
      const std::vector<unsigned short> &L = srcs[d];
      if (L.empty()) { continue; } // dummy slot (alignment): never stored
      if ((int)L.size() > 1 + PATCH_MAXEXTRA) { return false; }
      first[d] = L[0];
      if (L.size() > 1) { rec.push_back({{-(int)(L.size() - 1), L[0]}, &L}); }
   }
   // Order: most further sources first, then by destination.  The destinations of phase k are then exactly the first
   // n_k records, in the same order in every phase: entry i of every phase has the same destination, the thread that
   // owns index i (i mod #threads) performs all additions onto it in program order, and the phases need no barrier
   // between them.
   std::sort(rec.begin(), rec.end(), [](const auto &x, const auto &y) { return x.first < y.first; });
   rec.erase(std::unique(rec.begin(), rec.end(), [](const auto &x, const auto &y) { return x.first.second == y.first.second; }),
             rec.end());
   fold.assign(8, 0u);
   for (int k = 1; k <= PATCH_MAXEXTRA; k++)
   {
      unsigned n = 0;
      for (const auto &r : rec)
      {
         const std::vector<unsigned short> &L = *r.second;
         if ((int)L.size() <= k) { break; }
         fold.push_back((unsigned)L[0] | ((unsigned)L[k] << 16));
         n++;
      }
      fold[k - 1] = n;
   }
   return true;
}

// group (destination, staging index) pairs by destination; sources stay in ascending staging (= patch) order.
// Destinations with at most 4 sources go to the packed lists, the others to the CSR-like lists.
void group_by_dst(std::vector<std::pair<int, int>> &tup, PatchHost &H)
{
   std::stable_sort(tup.begin(), tup.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.first < b.first; });
   H.ptr.assign(1, 0);
   H.src.clear(); H.dst.clear(); H.src4.clear(); H.dst4.clear();
   size_t k = 0;
   while (k < tup.size())
   {
      size_t e = k;
      while (e < tup.size() && tup[e].first == tup[k].first) { e++; }
      if (e - k <= 4)
      {
         for (size_t q = 0; q < 4; q++) { H.src4.push_back(k + q < e ? tup[k + q].second : -1); }
         H.dst4.push_back(tup[k].first);
      }
      else
      {
         for (size_t q = k; q < e; q++) { H.src.push_back(tup[q].second); }
         H.ptr.push_back((int)H.src.size());
         H.dst.push_back(tup[k].first);
      }
      k = e;
   }
}
} // namespace

// Residual side: local rows per patch (interior / interface), row sources, interface reduction lists.
bool patch_build_y(Integrator &I, PatchHost &H)
{
   const int ne = I.ne, nvd = I.nvd, pe = I.pe, ld = I.pe + 1, np = (int)I.pdesc.size();
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
   std::vector<std::vector<unsigned char>> blobs(np);
   bool ok = true;
   parallel_for_p(np, 64, [&](long b, long e)
   {
      std::vector<int> vd, dofs, ifc, ylist;
      std::vector<std::vector<unsigned short>> srcs;
      std::vector<unsigned short> first;
      std::vector<unsigned> fold;
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
         srcs.assign(D.nrows, std::vector<unsigned short>());
         for (int l = 0; l < D.ne; l++)
         {
            build_vdofs(I, I.perm[lo + l], vd);
            for (int i = 0; i < nvd; i++)
            {
               const int v = vd[i];
               int lr;
               if (owner[v] != -2) { lr = (int)(std::lower_bound(R.begin(), R.begin() + D.nrow_int, v) - R.begin()); }
               else { lr = (int)(std::lower_bound(R.begin() + D.nrow_int, R.end(), v) - R.begin()); }
               srcs[lr].push_back((unsigned short)(i * ld + l));
            }
         }
         if (!pack_sources(srcs, D.nrows, first, fold)) { ok = false; continue; }
         D.nyfold = (int)fold.size();
         ylist.assign(R.begin(), R.begin() + D.nrow_int);
         BlobWriter W;
         W.section(first, first.size());
         W.section(ylist, ylist.size());
         W.section(fold, fold.size());
         blobs[p].swap(W.b);
      }
   });
   if (!ok) { return false; }
   long soff = 0, boff = 0;
   I.max_yblob = I.max_yg = I.max_yf = 0;
   for (int p = 0; p < np; p++)
   {
      PatchDesc &D = I.pdesc[p];
      D.ystage_off = (int)soff;
      D.yblob_off = (int)(boff / 16);
      D.yblob_bytes = (int)blobs[p].size();
      soff += D.nrows - D.nrow_int;
      boff += (long)blobs[p].size();
      I.prow_off[p + 1] = I.prow_off[p] + D.nrows;
      I.max_yblob = std::max(I.max_yblob, D.yblob_bytes);
      I.max_yg = std::max(I.max_yg, patch_yg_bytes(D));
      I.max_yf = std::max(I.max_yf, patch_al16(4 * D.nyfold));
   }
   H.stage_size = soff;
   H.blob.resize(std::max<long>(boff, 16));
   I.prows.resize(I.prow_off[np]);
   std::vector<std::pair<int, int>> tup; // (dof, stage index) in ascending patch order
   tup.reserve(soff);
   for (int p = 0; p < np; p++)
   {
      const PatchDesc &D = I.pdesc[p];
      std::copy(rows[p].begin(), rows[p].end(), I.prows.begin() + I.prow_off[p]);
      std::copy(blobs[p].begin(), blobs[p].end(), H.blob.begin() + (size_t)D.yblob_off * 16);
      for (int k = D.nrow_int; k < D.nrows; k++) { tup.emplace_back(rows[p][k], D.ystage_off + (k - D.nrow_int)); }
   }
   group_by_dst(tup, H);
   return true;
}

// Matrix side: slots of the patch, slot sources, runs of consecutive CSR positions, interface reduction lists.
// Slot order of a patch: [0,nint) rows interior to the patch, CSR order (runs of consecutive positions);
// [nint,nexc) entries of interface rows that only this patch contributes to (explicit CSR positions);
// [nexc,nslots) entries of interface rows shared with other patches.  The first two groups are final values
// and are written straight to the CSR array; only the third goes through staging.
bool patch_build_v(Integrator &I, PatchHost &H)
{
   const int nvd = I.nvd, pe = I.pe, ld = I.pe + 1, np = (int)I.pdesc.size();
   std::vector<std::vector<unsigned char>> blobs(np);
   std::vector<std::vector<long>> ifc_keys(np);  // (local row << 32 | column dof), sorted
   std::vector<std::vector<int>> ifc_gpos(np);   // CSR position of each key
   std::vector<std::vector<int>> shared_gpos(np); // CSR positions of the staged slots, staging order
   // pass A: interface entries of every patch
   parallel_for_p(np, 64, [&](long b, long e)
   {
      std::vector<int> vd;
      for (long p = b; p < e; p++)
      {
         const PatchDesc &D = I.pdesc[p];
         const int lo = (int)p * pe;
         const int *R = I.prows.data() + I.prow_off[p];
         std::vector<long> &keys = ifc_keys[p];
         keys.clear();
         for (int l = 0; l < D.ne; l++)
         {
            build_vdofs(I, I.perm[lo + l], vd);
            for (int i = 0; i < nvd; i++)
            {
               const int lr = (int)(std::lower_bound(R + D.nrow_int, R + D.nrows, vd[i]) - R);
               if (lr < D.nrows && R[lr] == vd[i])
               {
                  for (int j = 0; j < nvd; j++) { keys.push_back(((long)lr << 32) | (unsigned)vd[j]); }
               }
            }
         }
         std::sort(keys.begin(), keys.end());
         keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
         ifc_gpos[p].resize(keys.size());
         for (size_t k = 0; k < keys.size(); k++)
         {
            const int r = R[(int)(keys[k] >> 32)], c = (int)(keys[k] & 0xffffffff);
            const int *cb = I.colidx.data() + I.rowptr[r], *ce = I.colidx.data() + I.rowptr[r + 1];
            ifc_gpos[p][k] = (int)(std::lower_bound(cb, ce, c) - I.colidx.data());
         }
      }
   });
   // number of patches contributing to every CSR position of an interface row
   std::vector<unsigned char> cnt(I.colidx.size(), 0);
   for (int p = 0; p < np; p++)
   {
      for (int g : ifc_gpos[p]) { if (cnt[g] < 255) { cnt[g]++; } }
   }
   // pass B: slots, sources, blobs
   bool ok = true;
   parallel_for_p(np, 64, [&](long b, long e)
   {
      std::vector<int> vd, base, run_s, run_g, keyslot, xg, gpos, chunks, over, plist, glist;
      std::vector<unsigned short> isrc;
      std::vector<std::pair<int, int>> excl; // (CSR position, key index)
      std::vector<std::vector<unsigned short>> srcs;
      std::vector<unsigned short> first;
      std::vector<unsigned> fold;
      for (long p = b; p < e; p++)
      {
         PatchDesc &D = I.pdesc[p];
         const int lo = (int)p * pe;
         const int *R = I.prows.data() + I.prow_off[p];
         const std::vector<long> &keys = ifc_keys[p];
         // interior rows: slots follow the CSR rows; consecutive CSR positions merge into runs
         // A run that starts at an odd CSR position on an even slot (or vice versa) is shifted by one dummy slot (no
         // source, no store): the device writes aligned pairs of chunks (64 consecutive positions) with 16-byte stores.
         base.assign(D.nrows, 0);
         int s = 0;
         run_s.clear(); run_g.clear();
         gpos.clear();
         for (int lr = 0; lr < D.nrow_int; lr++)
         {
            const int r = R[lr];
            if (lr == 0 || R[lr - 1] + 1 != r)
            {
               if ((I.rowptr[r] ^ s) & 1) { gpos.push_back(-1); s++; }
               run_s.push_back(s); run_g.push_back(I.rowptr[r]);
            }
            base[lr] = s;
            for (int g = I.rowptr[r]; g < I.rowptr[r + 1]; g++) { gpos.push_back(g); }
            s += I.rowptr[r + 1] - I.rowptr[r];
         }
         // interface entries only this patch contributes to: final values, by CSR position
         excl.clear();
         keyslot.assign(keys.size(), -1);
         for (size_t k = 0; k < keys.size(); k++) { if (cnt[ifc_gpos[p][k]] == 1) { excl.emplace_back(ifc_gpos[p][k], (int)k); } }
         std::sort(excl.begin(), excl.end());
         D.nint = s;
         D.nruns = (int)run_s.size();
         run_s.push_back(s); // sentinel
         run_g.push_back(0);
         xg.clear();
         for (size_t k = 0; k < excl.size(); k++)
         {
            keyslot[excl[k].second] = s++;
            xg.push_back(excl[k].first);
         }
         D.nexc = s;
         // shared interface entries: staged
         shared_gpos[p].clear();
         for (size_t k = 0; k < keys.size(); k++)
         {
            if (keyslot[k] < 0)
            {
               keyslot[k] = s++;
               shared_gpos[p].push_back(ifc_gpos[p][k]);
            }
         }
         D.nslots = s;
         // element entries -> slots; a slot lists its sources in ascending element order
         srcs.assign(D.nslots, std::vector<unsigned short>());
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
                     slot = keyslot[std::lower_bound(keys.begin(), keys.end(), key) - keys.begin()];
                  }
                  const int lo_ = std::min(i, j), hi_ = std::max(i, j);
                  const int k = hi_ * (hi_ + 1) / 2 + lo_; // symidx (madb_kernels.cuh)
                  srcs[slot].push_back((unsigned short)(k * ld + l));
               }
            }
         }
         if (!pack_sources(srcs, D.nslots, first, fold)) { ok = false; continue; }
         D.nvfold = (int)fold.size();
         // CSR position of every directly written slot, packed per chunk of 32 slots:
         // {g0, g1 - split, split, n}: slots [0,split) of the chunk go to g0 + lane, the rest to g1 + (lane - split);
         // {g0, 0, 0, 64} + {0,0,0,0}: an (even, odd) pair of full chunks covering 64 consecutive positions, g0 even;
         // {0, 0, 0, 0}: nothing to write here (padding, second chunk of a pair, or an irregular chunk: more than one
         //               break or a dummy slot -> explicit positions over[] / sources isrc[], in chunk order)
         {
            gpos.resize(D.nexc, -1);
            for (int q = D.nint; q < D.nexc; q++) { gpos[q] = xg[q - D.nint]; }
            // chunk table padded to an even number of chunks (the device works on pairs);
            // d[3] = number of lanes of the chunk that store
            const int nchunk = ((D.nexc + 31) / 32 + 1) / 2 * 2;
            chunks.assign((size_t)4 * (nchunk + 2), 0); // + one all-zero pair: target of out-of-range pair indices
            over.clear();
            isrc.clear();
            for (int c = 0; c < nchunk; c++)
            {
               const int b0 = c * 32, n = std::max(0, std::min(32, D.nexc - b0));
               int split = n, breaks = 0;
               for (int q = 1; q < n; q++)
               {
                  if (gpos[b0 + q] != gpos[b0 + q - 1] + 1) { breaks++; if (breaks == 1) { split = q; } }
               }
               for (int q = 0; q < n; q++) { if (gpos[b0 + q] < 0) { breaks = 2; } } // dummy slot: explicit positions
               int *d = &chunks[(size_t)4 * c];
               if (n == 0) { continue; }
               if (breaks <= 1)
               {
                  d[0] = gpos[b0];
                  d[1] = (split < n) ? gpos[b0 + split] - split : 0;
                  d[2] = split;
                  d[3] = n;
               }
               else
               {
                  for (int q = 0; q < 32; q++)
                  {
                     over.push_back(q < n ? gpos[b0 + q] : -1);
                     isrc.push_back(q < n ? first[b0 + q] : 0);
                  }
               }
            }
            // pairs (2c, 2c+1) of full chunks covering 64 consecutive positions from an even one: {g0, 0, 0, 64}, {0,0,0,0}
            for (int c = 0; c + 1 < nchunk; c += 2)
            {
               int *d = &chunks[(size_t)4 * c];
               if (d[3] == 32 && d[2] == 32 && d[7] == 32 && d[6] == 32 && d[4] == d[0] + 32 && (d[0] & 1) == 0)
               {
                  d[1] = d[2] = 0;
                  d[3] = 64;
                  d[4] = d[5] = d[6] = d[7] = 0;
               }
            }
            while ((over.size() / 32) % 8 != 0) // pad the irregular list to a multiple of 8 chunks
            {
               for (int q = 0; q < 32; q++) { over.push_back(-1); isrc.push_back(0); }
            }
            // the two work lists of the device: aligned pairs {g0, first slot} and general chunks
            // {g0, g1 - split, split, n | first slot << 8}; each ends with one entry that stores nothing (the device
            // clamps out-of-range list indices onto it instead of branching)
            plist.clear();
            glist.clear();
            for (int c = 0; c < nchunk; c++)
            {
               const int *d = &chunks[(size_t)4 * c];
               if (d[3] == 64) { plist.push_back(d[0]); plist.push_back(32 * c); }
               else if (d[3] > 0)
               {
                  glist.push_back(d[0]); glist.push_back(d[1]); glist.push_back(d[2]);
                  glist.push_back(d[3] | ((32 * c) << 8));
               }
            }
            D.npair = (int)plist.size() / 2;
            D.ngen = (int)glist.size() / 4;
            plist.push_back(-1); plist.push_back(0);
            for (int q = 0; q < 4; q++) { glist.push_back(0); }
            D.nchunk = nchunk;
            D.nirr = (int)over.size() / 32;
            first.resize(std::max<size_t>(first.size(), (size_t)32 * (nchunk + 2)), 0); // padded so that chunk loads stay in bounds
         }
         D.nvsrc = (int)first.size();
         BlobWriter W;
         W.section(first, first.size());
         W.section(plist, plist.size());
         W.section(glist, glist.size());
         W.section(isrc, isrc.size());
         W.section(over, over.size());
         W.section(fold, fold.size());
         blobs[p].swap(W.b);
      }
   });
   if (!ok) { set_error("patch assembly: a matrix entry has more than 8 contributing elements in one patch"); return false; }
   long soff = 0, boff = 0;
   I.max_vblob = I.max_vg = I.max_vf = 0;
   for (int p = 0; p < np; p++)
   {
      PatchDesc &D = I.pdesc[p];
      D.stage_off = (int)soff;
      D.vblob_off = (int)(boff / 16);
      D.vblob_bytes = (int)blobs[p].size();
      soff += D.nslots - D.nexc;
      boff += (long)blobs[p].size();
      I.max_vblob = std::max(I.max_vblob, D.vblob_bytes);
      I.max_vg = std::max(I.max_vg, patch_vg_bytes(D));
      I.max_vf = std::max(I.max_vf, patch_al16(4 * D.nvfold));
   }
   if (boff / 16 >= 0x7fffffffL) { set_error("patch assembly: maps too large"); return false; }
   H.stage_size = soff;
   H.blob.resize(std::max<long>(boff, 16));
   std::vector<std::pair<int, int>> tup((size_t)soff);
   parallel_for_p(np, 64, [&](long b, long e)
   {
      for (long p = b; p < e; p++)
      {
         const PatchDesc &D = I.pdesc[p];
         std::copy(blobs[p].begin(), blobs[p].end(), H.blob.begin() + (size_t)D.vblob_off * 16);
         for (size_t k = 0; k < shared_gpos[p].size(); k++) { tup[(size_t)D.stage_off + k] = {shared_gpos[p][k], D.stage_off + (int)k}; }
      }
   });
   group_by_dst(tup, H);
   I.have_patch_vals = true;
   return true;
}


// ---------------------------------------------------------------------------------------------
// CSR-image kernel (k_patch_img): scatter maps, fold lists, runs.  See ImgDesc in madb_host.hpp.
// ---------------------------------------------------------------------------------------------
void img_schedule(int nvd, int tpe, int mirror_nd, std::vector<int> &vI, std::vector<int> &vJ, std::vector<int> &yI)
{
   if (tpe == 1)
   {
      // one thread per element: all entries of the upper triangle in packed order k = b (b + 1) / 2 + a
      const int nsym = nvd * (nvd + 1) / 2;
      vI.resize(nsym); vJ.resize(nsym); yI.resize(nvd);
      for (int b = 0; b < nvd; b++) { for (int a = 0; a <= b; a++) { vI[b * (b + 1) / 2 + a] = a; vJ[b * (b + 1) / 2 + a] = b; } }
      for (int i = 0; i < nvd; i++) { yI[i] = i; }
      return;
   }
   int nv = 0, ny = 0;
   switch (mirror_nd)
   {
      case 2: nv = sf2d_pair_keep_v<2>(nullptr, nullptr); ny = sf2d_pair_keep_y<2>(nullptr); break;
      case 3: nv = sf2d_pair_keep_v<3>(nullptr, nullptr); ny = sf2d_pair_keep_y<3>(nullptr); break;
      case 4: nv = sf2d_pair_keep_v<4>(nullptr, nullptr); ny = sf2d_pair_keep_y<4>(nullptr); break;
      case 5: nv = sf2d_pair_keep_v<5>(nullptr, nullptr); ny = sf2d_pair_keep_y<5>(nullptr); break;
      default: vI.clear(); vJ.clear(); yI.clear(); return;
   }
   vI.resize(nv); vJ.resize(nv); yI.resize(ny);
   switch (mirror_nd)
   {
      case 2: sf2d_pair_keep_v<2>(vI.data(), vJ.data()); sf2d_pair_keep_y<2>(yI.data()); break;
      case 3: sf2d_pair_keep_v<3>(vI.data(), vJ.data()); sf2d_pair_keep_y<3>(yI.data()); break;
      case 4: sf2d_pair_keep_v<4>(vI.data(), vJ.data()); sf2d_pair_keep_y<4>(yI.data()); break;
      default: sf2d_pair_keep_v<5>(vI.data(), vJ.data()); sf2d_pair_keep_y<5>(yI.data()); break;
   }
}

namespace
{
// fold list of one side: per destination the extras in ascending element order -> 8 phase counts + (dst | src << 16)
// words; records sorted by (number of extras descending, destination), so that entry i of every phase has the same
// destination and is handled by the same thread (no barrier between the phases)
bool pack_folds(std::vector<std::pair<int, std::vector<int>>> &rec, std::vector<int> &out)
{
   std::sort(rec.begin(), rec.end(), [](const auto &a, const auto &b)
   {
      if (a.second.size() != b.second.size()) { return a.second.size() > b.second.size(); }
      return a.first < b.first;
   });
   out.assign(8, 0);
   if (!rec.empty() && rec.front().second.size() > 8) { return false; }
   for (int k = 1; k <= 8; k++)
   {
      int n = 0;
      for (const auto &r : rec)
      {
         if ((int)r.second.size() < k) { break; }
         out.push_back((int)((unsigned)r.first | ((unsigned)r.second[k - 1] << 16)));
         n++;
      }
      out[k - 1] = n;
   }
   return true;
}
void pad4(std::vector<int> &v) { while (v.size() % 4) { v.push_back(0); } }
} // namespace

bool patch_build_img(Integrator &I, PatchHost &H, ImgHost &IH)
{
   const int nvd = I.nvd, pe = I.pe, np = (int)I.pdesc.size(), tpe = I.ops.img_tpe, mnd = I.ops.img_mirror_nd;
   std::vector<int> eI, eJ, eY;
   img_schedule(nvd, tpe, mnd, eI, eJ, eY);
   const int nev = (int)eI.size(), ney = (int)eY.size();
   if (nev == 0) { set_error("patch_build_img: no emission schedule for this configuration"); return false; }
   auto mir = [&](int i, int h) { return (h && tpe == 2) ? (mnd - 1 - i / mnd) * mnd + i % mnd : i; };
   IH.nev = nev; IH.ney = ney;
   IH.desc.assign(np, ImgDesc());
   std::vector<std::vector<unsigned char>> mblobs(np);
   std::vector<std::vector<int>> gdatas(np);
   std::vector<std::vector<int>> shared_gpos(np);
   // interface entries of every patch, number of contributing patches per CSR position (as in patch_build_v)
   std::vector<std::vector<long>> ifc_keys(np);
   std::vector<std::vector<int>> ifc_gpos(np);
   parallel_for_p(np, 64, [&](long b, long e)
   {
      std::vector<int> vd;
      for (long p = b; p < e; p++)
      {
         const PatchDesc &D = I.pdesc[p];
         const int lo = (int)p * pe;
         const int *R = I.prows.data() + I.prow_off[p];
         std::vector<long> &keys = ifc_keys[p];
         keys.clear();
         for (int l = 0; l < D.ne; l++)
         {
            build_vdofs(I, I.perm[lo + l], vd);
            for (int i = 0; i < nvd; i++)
            {
               const int lr = (int)(std::lower_bound(R + D.nrow_int, R + D.nrows, vd[i]) - R);
               if (lr < D.nrows && R[lr] == vd[i])
               {
                  for (int j = 0; j < nvd; j++) { keys.push_back(((long)lr << 32) | (unsigned)vd[j]); }
               }
            }
         }
         std::sort(keys.begin(), keys.end());
         keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
         ifc_gpos[p].resize(keys.size());
         for (size_t k = 0; k < keys.size(); k++)
         {
            const int r = R[(int)(keys[k] >> 32)], c = (int)(keys[k] & 0xffffffff);
            const int *cb = I.colidx.data() + I.rowptr[r], *ce = I.colidx.data() + I.rowptr[r + 1];
            ifc_gpos[p][k] = (int)(std::lower_bound(cb, ce, c) - I.colidx.data());
         }
      }
   });
   std::vector<unsigned char> cnt(I.colidx.size(), 0);
   for (int p = 0; p < np; p++) { for (int g : ifc_gpos[p]) { if (cnt[g] < 255) { cnt[g]++; } } }

   bool ok = true;
   parallel_for_p(np, 64, [&](long b, long e)
   {
      std::vector<int> vd, base, keyslot, vfold, yfold, runs, xgpos, xslot, scnt, ycnt;
      std::vector<std::pair<int, int>> excl;
      std::vector<unsigned> word; // [l][k] scatter word of sym entry k of element l
      std::vector<unsigned short> yword;
      std::vector<std::pair<int, std::vector<int>>> vrec, yrec;
      std::vector<int> recof; // slot -> record index (or -1)
      for (long p = b; p < e; p++)
      {
         const PatchDesc &D = I.pdesc[p];
         ImgDesc &G = IH.desc[p];
         std::memset(&G, 0, sizeof(G));
         const int lo = (int)p * pe;
         const int *R = I.prows.data() + I.prow_off[p];
         const std::vector<long> &keys = ifc_keys[p];
         // ---- image layout ------------------------------------------------------------------------------
         base.assign(D.nrows, 0);
         runs.clear();
         int s = 0;
         for (int lr = 0; lr < D.nrow_int; lr++)
         {
            const int r = R[lr];
            if (lr == 0 || R[lr - 1] + 1 != r)
            {
               if ((I.rowptr[r] ^ s) & 1) { s++; } // same parity of image slot and CSR position
               runs.push_back(s); runs.push_back(I.rowptr[r]); runs.push_back(0); runs.push_back(0);
            }
            base[lr] = s;
            s += I.rowptr[r + 1] - I.rowptr[r];
            runs[runs.size() - 2] += I.rowptr[r + 1] - I.rowptr[r];
         }
         excl.clear();
         keyslot.assign(keys.size(), -1);
         for (size_t k = 0; k < keys.size(); k++) { if (cnt[ifc_gpos[p][k]] == 1) { excl.emplace_back(ifc_gpos[p][k], (int)k); } }
         std::sort(excl.begin(), excl.end());
         xgpos.clear(); xslot.clear();
         for (size_t k = 0; k < excl.size(); k++)
         {
            keyslot[excl[k].second] = s;
            xgpos.push_back(excl[k].first);
            xslot.push_back(s);
            s++;
         }
         if (s & 1) { s++; }
         const int sh0 = s;
         shared_gpos[p].clear();
         for (size_t k = 0; k < keys.size(); k++)
         {
            if (keyslot[k] < 0)
            {
               keyslot[k] = s++;
               shared_gpos[p].push_back(ifc_gpos[p][k]);
            }
         }
         if (s & 1) { s++; } // pad slot: copied to a staging position nobody reads
         const int nsh = s - sh0;
         const int trash = s;
         s += 2;
         int xnext = s; // extras
         // ---- sources: first source -> the slot, further sources -> extras + fold ----------------------------
         const int nsym = nvd * (nvd + 1) / 2;
         word.assign((size_t)pe * nsym, (unsigned)trash | ((unsigned)trash << 16));
         scnt.assign(xnext, 0);
         recof.assign(xnext, -1);
         vrec.clear();
         const int ytrash = D.nrows;
         int ynext = D.nrows + 1;
         yword.assign((size_t)pe * nvd, (unsigned short)ytrash);
         ycnt.assign(D.nrows, 0);
         yrec.clear();
         std::vector<int> yrecof(D.nrows, -1), lrow(nvd), rbeg(nvd);
         for (int l = 0; l < D.ne; l++)
         {
            build_vdofs(I, I.perm[lo + l], vd);
            for (int i = 0; i < nvd; i++)
            {
               int lr = (int)(std::lower_bound(R, R + D.nrow_int, vd[i]) - R);
               const bool interior = lr < D.nrow_int && R[lr] == vd[i];
               if (!interior) { lr = (int)(std::lower_bound(R + D.nrow_int, R + D.nrows, vd[i]) - R); }
               lrow[i] = interior ? lr : -1 - lr;
               // element vector
               if (ycnt[lr]++ == 0) { yword[(size_t)l * nvd + i] = (unsigned short)lr; }
               else
               {
                  if (yrecof[lr] < 0) { yrecof[lr] = (int)yrec.size(); yrec.push_back({lr, {}}); }
                  yrec[yrecof[lr]].second.push_back(ynext);
                  yword[(size_t)l * nvd + i] = (unsigned short)ynext++;
               }
            }
            auto slot_of = [&](int i, int j)
            {
               if (lrow[i] >= 0)
               {
                  const int r = vd[i];
                  const int *cb = I.colidx.data() + I.rowptr[r], *ce = I.colidx.data() + I.rowptr[r + 1];
                  return base[lrow[i]] + (int)(std::lower_bound(cb, ce, vd[j]) - cb);
               }
               const long key = ((long)(-1 - lrow[i]) << 32) | (unsigned)vd[j];
               return keyslot[std::lower_bound(keys.begin(), keys.end(), key) - keys.begin()];
            };
            for (int jb = 0; jb < nvd; jb++)
            {
               for (int ia = 0; ia <= jb; ia++)
               {
                  const int k = jb * (jb + 1) / 2 + ia;
                  const int s1 = slot_of(ia, jb), s2 = slot_of(jb, ia);
                  const int rank = scnt[s1]++;
                  if (s2 != s1) { scnt[s2]++; }
                  if (rank == 0) { word[(size_t)l * nsym + k] = (unsigned)s1 | ((unsigned)s2 << 16); }
                  else
                  {
                     const int x = xnext++;
                     word[(size_t)l * nsym + k] = (unsigned)x | ((unsigned)x << 16);
                     for (int sdst : {s1, s2})
                     {
                        if (recof[sdst] < 0) { recof[sdst] = (int)vrec.size(); vrec.push_back({sdst, {}}); }
                        vrec[recof[sdst]].second.push_back(x);
                        if (s2 == s1) { break; }
                     }
                  }
               }
            }
         }
         if (xnext > 65535 || ynext > 65535) { ok = false; continue; }
         if (!pack_folds(vrec, vfold) || !pack_folds(yrec, yfold)) { ok = false; continue; }
         // ---- blobs ------------------------------------------------------------------------------------------
         std::vector<int> &gd = gdatas[p];
         gd.clear();
         gd.insert(gd.end(), vfold.begin(), vfold.end()); pad4(gd);
         gd.insert(gd.end(), yfold.begin(), yfold.end()); pad4(gd);
         gd.insert(gd.end(), runs.begin(), runs.end()); pad4(gd);
         gd.insert(gd.end(), xgpos.begin(), xgpos.end()); pad4(gd);
         for (size_t k = 0; k < xslot.size(); k += 2)
         {
            gd.push_back((int)((unsigned)xslot[k] | ((k + 1 < xslot.size() ? (unsigned)xslot[k + 1] : 0u) << 16)));
         }
         pad4(gd);
         gd.insert(gd.end(), R, R + D.nrow_int); pad4(gd);
         // one blob per patch, fetched with one bulk copy: descriptor | scatter maps | lists
         std::vector<unsigned char> &mb = mblobs[p];
         const size_t vbytes = (size_t)nev * pe * tpe * 4, ybytes = (size_t)ney * pe * tpe * 2;
         const size_t o_vmap = (sizeof(ImgDesc) + 15) & ~(size_t)15;
         const size_t o_ymap = (o_vmap + vbytes + 15) & ~(size_t)15;
         const size_t o_lists = (o_ymap + ybytes + 15) & ~(size_t)15;
         mb.assign((o_lists + gd.size() * 4 + 15) & ~(size_t)15, 0);
         std::memcpy(mb.data() + o_lists, gd.data(), gd.size() * 4);
         unsigned *vm = (unsigned *)(mb.data() + o_vmap);
         unsigned short *ym = (unsigned short *)(mb.data() + o_ymap);
         for (int ee = 0; ee < nev; ee++)
         {
            for (int l = 0; l < pe; l++)
            {
               for (int h = 0; h < tpe; h++)
               {
                  const int a = mir(eI[ee], h), bb = mir(eJ[ee], h);
                  const int lo_ = std::min(a, bb), hi_ = std::max(a, bb);
                  vm[((size_t)ee * pe + l) * tpe + h] = word[(size_t)l * nsym + hi_ * (hi_ + 1) / 2 + lo_];
               }
            }
         }
         for (int ee = 0; ee < ney; ee++)
         {
            for (int l = 0; l < pe; l++)
            {
               for (int h = 0; h < tpe; h++) { ym[((size_t)ee * pe + l) * tpe + h] = yword[(size_t)l * nvd + mir(eY[ee], h)]; }
            }
         }
         G.o_vmap = (int)o_vmap;
         G.o_lists = (int)o_lists;
         G.ne = D.ne; G.nrows = D.nrows; G.nrow_int = D.nrow_int; G.ystage_off = D.ystage_off;
         G.mblob_bytes = (int)mb.size();
         G.o_ymap = (int)o_ymap;
         G.nvfold = (int)vfold.size(); G.nyfold = (int)yfold.size();
         G.nruns = (int)runs.size() / 4; G.nexcl = (int)xgpos.size();
         G.sh0 = sh0; G.nsh = nsh;
         G.nvslots = xnext; G.nyslots = ynext;
      }
   });
   if (!ok) { set_error("patch assembly (CSR image): a patch does not fit the 16-bit slot indices / 8 fold phases"); return false; }
   long soff = 0, boff = 0;
   IH.max_vslots = IH.max_yslots = IH.max_mblob = 0;
   for (int p = 0; p < np; p++)
   {
      ImgDesc &G = IH.desc[p];
      G.stage_off = (int)soff;
      G.mblob_off = (int)(boff / 16);
      soff += G.nsh;
      boff += (long)mblobs[p].size();
      IH.max_vslots = std::max(IH.max_vslots, G.nvslots);
      IH.max_yslots = std::max(IH.max_yslots, G.nyslots);
      IH.max_mblob = std::max(IH.max_mblob, G.mblob_bytes);
   }
   if (boff / 16 >= 0x7fffffffL) { set_error("patch assembly (CSR image): maps too large"); return false; }
   H.stage_size = soff;
   IH.mblob.resize(std::max<long>(boff, 16));
   std::vector<std::pair<int, int>> tup;
   for (int p = 0; p < np; p++)
   {
      const ImgDesc &G = IH.desc[p];
      std::copy(mblobs[p].begin(), mblobs[p].end(), IH.mblob.begin() + (size_t)G.mblob_off * 16);
      std::memcpy(IH.mblob.data() + (size_t)G.mblob_off * 16, &G, sizeof(ImgDesc)); // the blob starts with its descriptor
      for (size_t k = 0; k < shared_gpos[p].size(); k++) { tup.emplace_back(shared_gpos[p][k], G.stage_off + (int)k); }
   }
   group_by_dst(tup, H);
   I.img_map_bytes = boff;
   I.have_patch_vals = true;
   return true;
}

} // namespace madb

// ---------------------------------------------------------------------------------------------
// Host-only self test of the patch maps (no CUDA): builds the maps for one H1 field, assembles
// integer-valued element vectors / symmetric element matrices (exact in FP64, so the summation
// order does not matter) once directly and once by emulating what the kernels do with the maps
// (stage, fold phases, first-source gather, chunk descriptors, irregular chunks, staging,
// interface reduction), and returns the largest difference.  Used by the CPU test-suite.
// ---------------------------------------------------------------------------------------------
namespace madb
{

int patch_selftest(Mesh &mesh, Space &space, double *max_err, long *stats)
{
   Integrator I;
   I.mesh = &mesh;
   FieldDesc fd;
   fd.space = &space;
   fd.mode = 4; // GRAD
   fd.role = 0;
   I.fields.push_back(fd);
   I.ne = mesh.ne;
   I.nvd = space.nd_el() * space.vdim;
   I.ndof_all = I.nvd;
   I.ntotal = space.vsize();
   I.goff.assign(1, 0);
   if (!patch_eligible(I.nvd)) { set_error("patch_selftest: element matrix too large for the patch scheme"); return 1; }
   I.pe = patch_pe(I.nvd);
   I.use_patches = true;
   patch_order(I);
   if (I.pdesc.empty()) { set_error("patch_selftest: patch_order failed"); return 1; }
   const int pe = I.pe, ld = pe + 1, nvd = I.nvd, nsym = nvd * (nvd + 1) / 2, np = (int)I.pdesc.size();
   I.stride = np * pe;
   PatchHost HY, HV;
   if (!patch_build_y(I, HY)) { set_error("patch_selftest: patch_build_y failed"); return 1; }
   build_pattern(I);
   if (!I.have_pattern) { return 1; }
   if (!patch_build_v(I, HV)) { return 1; }
   const long N = I.ntotal, nnz = (long)I.colidx.size();
   auto rval = [](int e, int i) { return (double)((e * 7 + i * 3) % 11 - 5); };
   auto aval = [](int e, int a, int b) { const int lo = std::min(a, b), hi = std::max(a, b); return (double)((e * 5 + lo * 13 + hi * 17) % 23 - 11); };
   // direct assembly
   std::vector<double> y_ref(N, 0.0), v_ref(nnz, 0.0), y(N, -777.0), v(nnz, -777.0);
   std::vector<int> vd;
   for (int e = 0; e < I.ne; e++)
   {
      build_vdofs(I, e, vd);
      for (int i = 0; i < nvd; i++)
      {
         y_ref[vd[i]] += rval(e, i);
         const int *cb = I.colidx.data() + I.rowptr[vd[i]], *ce = I.colidx.data() + I.rowptr[vd[i] + 1];
         for (int j = 0; j < nvd; j++) { v_ref[std::lower_bound(cb, ce, vd[j]) - I.colidx.data()] += aval(e, i, j); }
      }
   }
   // emulation of the kernels
   std::vector<double> ystage(std::max<long>(HY.stage_size, 1), -555.0), vstage(std::max<long>(HV.stage_size, 1), -555.0);
   long npaired = 0; // CSR entries written through the 16-byte path
   std::vector<double> sR((size_t)nvd * ld), sA((size_t)nsym * ld);
   for (int p = 0; p < np; p++)
   {
      const PatchDesc &D = I.pdesc[p];
      std::fill(sR.begin(), sR.end(), 1e300);
      std::fill(sA.begin(), sA.end(), 1e300);
      for (int l = 0; l < D.ne; l++)
      {
         const int e = I.perm[p * pe + l];
         for (int i = 0; i < nvd; i++) { sR[(size_t)i * ld + l] = rval(e, i); }
         for (int b = 0; b < nvd; b++) { for (int a = 0; a <= b; a++) { sA[(size_t)(b * (b + 1) / 2 + a) * ld + l] = aval(e, a, b); } }
      }
      // y side
      {
         const unsigned char *yb = HY.blob.data() + (size_t)D.yblob_off * 16;
         const unsigned short *ysrc = (const unsigned short *)yb;
         const int *ylist = (const int *)(yb + patch_al16(2 * D.nrows));
         const unsigned *yfold = (const unsigned *)(yb + patch_yg_bytes(D));
         int base = 8;
         for (int ph = 0; ph < 8; ph++)
         {
            const int n = (int)yfold[ph];
            for (int i = 0; i < n; i++) { const unsigned w = yfold[base + i]; sR[w & 0xffffu] += sR[w >> 16]; }
            base += n;
         }
         for (int lr = 0; lr < D.nrows; lr++)
         {
            const double val = sR[ysrc[lr]];
            if (lr < D.nrow_int) { y[ylist[lr]] = val; }
            else { ystage[D.ystage_off + (lr - D.nrow_int)] = val; }
         }
      }
      // matrix side
      {
         const unsigned char *vb = HV.blob.data() + (size_t)D.vblob_off * 16;
         const unsigned short *vsrc = (const unsigned short *)vb;
         const unsigned *vfold = (const unsigned *)(vb + patch_vg_bytes(D));
         const int *plist = (const int *)(vb + patch_al16(2 * D.nvsrc));
         const int *glist = (const int *)((const unsigned char *)plist + patch_al16(8 * (D.npair + 1)));
         const unsigned short *isrc = (const unsigned short *)((const unsigned char *)glist + patch_al16(16 * (D.ngen + 1)));
         const int *over = (const int *)((const unsigned char *)isrc + patch_al16(64 * D.nirr));
         int base = 8;
         for (int ph = 0; ph < 8; ph++)
         {
            const int n = (int)vfold[ph];
            for (int i = 0; i < n; i++) { const unsigned w = vfold[base + i]; sA[w & 0xffffu] += sA[w >> 16]; }
            base += n;
         }
         // aligned pairs of chunks: lane writes the slots 2 lane, 2 lane + 1 (one 16-byte store)
         for (int k = 0; k <= D.npair; k++)
         {
            const int g0 = plist[2 * k], s0 = plist[2 * k + 1];
            if (g0 < 0) { continue; }
            npaired += 64;
            for (int q = 0; q < 64; q++) { v[g0 + q] = sA[vsrc[s0 + q]]; }
         }
         for (int k = 0; k <= D.ngen; k++)
         {
            const int *d = glist + 4 * k;
            const int n = d[3] & 0xff, s0 = d[3] >> 8;
            for (int lane = 0; lane < n; lane++) { v[((lane < d[2]) ? d[0] : d[1]) + lane] = sA[vsrc[s0 + lane]]; }
         }
         for (int k = 0; k < D.nirr * 32; k++) { if (over[k] >= 0) { v[over[k]] = sA[isrc[k]]; } }
         for (int s = D.nexc; s < D.nslots; s++) { vstage[D.stage_off + (s - D.nexc)] = sA[vsrc[s]]; }
      }
   }
   auto reduce = [](const PatchHost &H, const std::vector<double> &stage, std::vector<double> &out)
   {
      for (size_t i = 0; i < H.dst4.size(); i++)
      {
         double s = 0.0;
         for (int k = 0; k < 4; k++) { if (H.src4[4 * i + k] >= 0) { s += stage[H.src4[4 * i + k]]; } }
         out[H.dst4[i]] = s;
      }
      for (size_t i = 0; i < H.dst.size(); i++)
      {
         double s = 0.0;
         for (int k = H.ptr[i]; k < H.ptr[i + 1]; k++) { s += stage[H.src[k]]; }
         out[H.dst[i]] = s;
      }
   };
   reduce(HY, ystage, y);
   reduce(HV, vstage, v);
   double err = 0.0;
   for (long i = 0; i < N; i++) { err = std::max(err, std::fabs(y[i] - y_ref[i])); }
   for (long i = 0; i < nnz; i++) { err = std::max(err, std::fabs(v[i] - v_ref[i])); }
   *max_err = err;
   if (stats)
   {
      stats[0] = np;
      stats[1] = (long)HY.dst4.size() + (long)HY.dst.size();
      stats[2] = (long)HV.dst4.size() + (long)HV.dst.size();
      stats[3] = HV.stage_size;
      stats[4] = I.max_vblob;
      stats[5] = nnz;
      stats[6] = npaired;
      stats[7] = 0;
   }
   return 0;
}


// Host-only emulation of k_patch_img (madb_patch_img.cuh) on integer-valued element data: scatter through the maps in
// the emission order of the element threads (tpe = 2: the thread pair with the mirrored half-element), fold, runs with
// the parity peeling of the bulk copies, exclusive / shared interface entries, interface reduction.
int patch_selftest_img(Mesh &mesh, Space &space, int tpe, double *max_err, long *stats)
{
   Integrator I;
   I.mesh = &mesh;
   FieldDesc fd;
   fd.space = &space;
   fd.mode = 4; // GRAD
   fd.role = 0;
   I.fields.push_back(fd);
   I.ne = mesh.ne;
   I.nvd = space.nd_el() * space.vdim;
   I.ndof_all = I.nvd;
   I.ntotal = space.vsize();
   I.goff.assign(1, 0);
   if (patch_pe(I.nvd) != PATCH_PE) { set_error("patch_selftest_img: element matrix too large for the CSR-image kernel"); return 1; }
   if (tpe != 1 && !(tpe == 2 && space.vdim == 1 && mesh.dim == 2)) { set_error("patch_selftest_img: thread pairs need a scalar 2-D space"); return 1; }
   I.ops.img_tpe = tpe;
   I.ops.img_mirror_nd = tpe == 2 ? space.order + 1 : 0;
   I.pe = 128 / tpe;
   I.use_patches = true;
   patch_order(I);
   if (I.pdesc.empty()) { set_error("patch_selftest_img: patch_order failed"); return 1; }
   const int pe = I.pe, nvd = I.nvd, np = (int)I.pdesc.size();
   I.stride = np * pe;
   PatchHost HY, HV;
   ImgHost IH;
   if (!patch_build_y(I, HY)) { set_error("patch_selftest_img: patch_build_y failed"); return 1; }
   build_pattern(I);
   if (!I.have_pattern) { return 1; }
   if (!patch_build_img(I, HV, IH)) { return 1; }
   std::vector<int> eI, eJ, eY;
   img_schedule(nvd, tpe, I.ops.img_mirror_nd, eI, eJ, eY);
   const int nev = (int)eI.size(), ney = (int)eY.size(), mnd = I.ops.img_mirror_nd;
   auto mir = [&](int i, int h) { return (h && tpe == 2) ? (mnd - 1 - i / mnd) * mnd + i % mnd : i; };
   const long N = I.ntotal, nnz = (long)I.colidx.size();
   auto rval = [](int e, int i) { return (double)((e * 7 + i * 3) % 11 - 5); };
   auto aval = [](int e, int a, int b) { const int lo = std::min(a, b), hi = std::max(a, b); return (double)((e * 5 + lo * 13 + hi * 17) % 23 - 11); };
   std::vector<double> y_ref(N, 0.0), v_ref(nnz, 0.0), y(N, -777.0), v(nnz, -777.0);
   std::vector<int> vd;
   for (int e = 0; e < I.ne; e++)
   {
      build_vdofs(I, e, vd);
      for (int i = 0; i < nvd; i++)
      {
         y_ref[vd[i]] += rval(e, i);
         const int *cb = I.colidx.data() + I.rowptr[vd[i]], *ce = I.colidx.data() + I.rowptr[vd[i] + 1];
         for (int j = 0; j < nvd; j++) { v_ref[std::lower_bound(cb, ce, vd[j]) - I.colidx.data()] += aval(e, i, j); }
      }
   }
   std::vector<double> ystage(std::max<long>(HY.stage_size, 1), -555.0), vstage(std::max<long>(HV.stage_size, 2), -555.0);
   std::vector<double> img(IH.max_vslots), yimg(IH.max_yslots);
   long nbulk = 0, nwritten = 0;
   for (int p = 0; p < np; p++)
   {
      const ImgDesc &D = IH.desc[p];
      std::fill(img.begin(), img.end(), 1e300);
      std::fill(yimg.begin(), yimg.end(), 1e300);
      const unsigned char *blob = IH.mblob.data() + (size_t)D.mblob_off * 16;
      if (std::memcmp(blob, &D, sizeof(ImgDesc)) != 0) { set_error("patch_selftest_img: blob descriptor"); return 1; }
      const unsigned *vm = (const unsigned *)(blob + D.o_vmap);
      const unsigned short *ym = (const unsigned short *)(blob + D.o_ymap);
      for (int l = 0; l < D.ne; l++)
      {
         const int e = I.perm[p * pe + l];
         for (int h = 0; h < tpe; h++)
         {
            for (int ee = 0; ee < ney; ee++) { yimg[ym[((size_t)ee * pe + l) * tpe + h]] = rval(e, mir(eY[ee], h)); }
            for (int ee = 0; ee < nev; ee++)
            {
               const unsigned w = vm[((size_t)ee * pe + l) * tpe + h];
               const double val = aval(e, mir(eI[ee], h), mir(eJ[ee], h));
               img[w & 0xffffu] = val;
               img[w >> 16] = val;
            }
         }
      }
      const int *gd = (const int *)(blob + D.o_lists);
      auto fold = [](const int *f, std::vector<double> &a)
      {
         int base = 8;
         for (int ph = 0; ph < 8; ph++)
         {
            const int n = f[ph];
            for (int i = 0; i < n; i++) { const unsigned w = (unsigned)f[base + i]; a[w & 0xffffu] += a[w >> 16]; }
            base += n;
         }
      };
      fold(gd, img);
      fold(gd + img_al4(D.nvfold), yimg);
      const int *runs = gd + img_al4(D.nvfold) + img_al4(D.nyfold);
      const int *xg = runs + 4 * D.nruns;
      const unsigned short *xs = (const unsigned short *)(xg + img_al4(D.nexcl));
      const int *ylist = xg + img_al4(D.nexcl) + img_al4((D.nexcl + 1) / 2);
      for (int r = 0; r < D.nruns; r++)
      {
         int so = runs[4 * r], g0 = runs[4 * r + 1], n = runs[4 * r + 2];
         if ((so ^ g0) & 1) { set_error("patch_selftest_img: run parity"); return 1; }
         nwritten += n;
         if (g0 & 1) { v[g0] = img[so]; so++; g0++; n--; }
         if (n & 1) { v[g0 + n - 1] = img[so + n - 1]; n--; }
         for (int j = 0; j < n; j++) { v[g0 + j] = img[so + j]; }
         nbulk += n;
      }
      if ((D.sh0 & 1) || (D.nsh & 1) || (D.stage_off & 1)) { set_error("patch_selftest_img: staging alignment"); return 1; }
      for (int k = 0; k < D.nsh; k++) { vstage[D.stage_off + k] = img[D.sh0 + k]; }
      for (int q = 0; q < D.nexcl; q++) { v[xg[q]] = img[xs[q]]; }
      for (int lr = 0; lr < D.nrows; lr++)
      {
         if (lr < D.nrow_int) { y[ylist[lr]] = yimg[lr]; }
         else { ystage[D.ystage_off + (lr - D.nrow_int)] = yimg[lr]; }
      }
   }
   auto reduce = [](const PatchHost &H, const std::vector<double> &stage, std::vector<double> &out)
   {
      for (size_t i = 0; i < H.dst4.size(); i++)
      {
         double s = 0.0;
         for (int k = 0; k < 4; k++) { if (H.src4[4 * i + k] >= 0) { s += stage[H.src4[4 * i + k]]; } }
         out[H.dst4[i]] = s;
      }
      for (size_t i = 0; i < H.dst.size(); i++)
      {
         double s = 0.0;
         for (int k = H.ptr[i]; k < H.ptr[i + 1]; k++) { s += stage[H.src[k]]; }
         out[H.dst[i]] = s;
      }
   };
   reduce(HY, ystage, y);
   reduce(HV, vstage, v);
   double err = 0.0;
   for (long i = 0; i < N; i++) { err = std::max(err, std::fabs(y[i] - y_ref[i])); }
   for (long i = 0; i < nnz; i++) { err = std::max(err, std::fabs(v[i] - v_ref[i])); }
   *max_err = err;
   if (stats)
   {
      stats[0] = np;
      stats[1] = (long)HY.dst4.size() + (long)HY.dst.size();
      stats[2] = (long)HV.dst4.size() + (long)HV.dst.size();
      stats[3] = HV.stage_size;
      stats[4] = (long)IH.max_vslots * 8 + (long)IH.max_yslots * 8 + IH.max_mblob; // shared memory per work group
      stats[5] = nnz;
      stats[6] = nbulk;           // CSR entries written by bulk copies
      stats[7] = I.img_map_bytes; // bytes of maps + lists read per assembly
   }
   return 0;
}

} // namespace madb
