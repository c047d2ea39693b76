// madb_pair_schedule.hpp -- which element-vector / matrix entries a thread of the pair kernel keeps, in emission order
// (shared by the device code, madb_sf2d_pair.cuh, and the host-side map builder, madb_patch.cpp).
#pragma once
#if defined(__CUDACC__)
#define MADB_SCHED_HD __host__ __device__ inline
#else
#define MADB_SCHED_HD inline
#endif

namespace madb
{

/// Walk of the kept matrix entries: for emission index e the LOCAL (I, J) the thread keeps (I = i2*ND + i1).
/// Blocks (j1, i1 <= j1) in the order of the matrix phase; inside a block the (i2, j2) with (i2, j2) <= mirror in
/// lexicographic order are kept.  Returns the number of kept entries; fills I[], J[] when non-null.
template <int ND> MADB_SCHED_HD constexpr int sf2d_pair_keep_v(int *Iout, int *Jout)
{
   int e = 0;
   for (int j1 = 0; j1 < ND; j1++)
   {
      for (int i1 = 0; i1 <= j1; i1++)
      {
         for (int i2 = 0; i2 < ND; i2++)
         {
            for (int j2 = 0; j2 < ND; j2++)
            {
               if (i1 == j1 && i2 > j2) { continue; } // diagonal block: upper triangle only
               // partner entry under the mirror: (ND-1-i2, ND-1-j2); in a diagonal block it is stored with sorted indices
               int mi = ND - 1 - i2, mj = ND - 1 - j2;
               if (i1 == j1 && mi > mj) { const int t = mi; mi = mj; mj = t; }
               const bool keep = (i2 < mi) || (i2 == mi && j2 <= mj);
               if (!keep) { continue; }
               if (Iout) { Iout[e] = i2 * ND + i1; Jout[e] = j2 * ND + j1; }
               e++;
            }
         }
      }
   }
   return e;
}
/// Walk of the kept element-vector entries: rows i2 <= ND-1-i2.
template <int ND> MADB_SCHED_HD constexpr int sf2d_pair_keep_y(int *Iout)
{
   int e = 0;
   for (int i2 = 0; 2 * i2 <= ND - 1; i2++)
   {
      for (int i1 = 0; i1 < ND; i1++)
      {
         if (Iout) { Iout[e] = i2 * ND + i1; }
         e++;
      }
   }
   return e;
}


} // namespace madb
