// madb_kernels3d.cuh -- sum-factorised AD residual, matrix-free Jacobian action and
// energy for tensor-product hexahedra (config 3: H1 order 3, 5^3 Gauss points).
//
// Same mathematics as madb_kernels.cuh (reference: src/ad_intg.hpp:202-257 for the
// residual; the action y = J(x) v has no reference counterpart -- the assembled
// Jacobian of config 3 would be 30 GB), but the dof<->quadrature contractions
// are done one direction at a time (O(p^4) instead of O(p^6) per element):
//   block  = (NQ, NQ, NEB) threads: one (qx,qy) column per thread, NEB elements per CTA
//   x-step : u[ix,iy,iz]      -> A,C[qx,iy,iz]        (shared memory)
//   y-step : A,C              -> BB,GB,BG[qx,qy,iz]   (registers, per thread)
//   z-step : loop over qz: reference gradient, trilinear Jacobian from the 8 vertices,
//            one AD pass of the functional, pull-back, accumulate the reverse z-step
//   reverse y/x steps through shared memory, coloured deterministic scatter.
#pragma once
#include "madb_ad.cuh"
#include "madb_host.hpp"
#include "madb_kernels.cuh"

namespace madb
{

template <class Func, int ND, int NQ> struct Sf3Args
{
   static_assert(Func::N_INPUT == 3, "sum-factorised 3-D kernel: scalar space with ADEval::GRAD");
   static_assert(Func::N_QPRM == 0, "sum-factorised 3-D kernel: no per-point parameters yet");
   int begin, end;
   const int *e2n;       // [t][8]
   const double *coords; // [nnodes][3]
   const int *vmap;      // [t][ND^3]  (bit 31: first touch)
   const double *x, *v;
   double *y, *energy;
   double fparams[Func::N_PARAM > 0 ? Func::N_PARAM : 1];
   double B[NQ][ND], G[NQ][ND], xq[NQ], wq[NQ];
};

// Minimum CTAs per SM asked of ptxas (a register cap).  Measured on config 3 (104^3 Q3, 125-thread CTAs): residual 6.86 / 4.94 /
// 4.15 / 3.86 ms and action 7.82 / 5.79 / 5.03 / 5.24 ms for 1 / 3 / 4 / 5 CTAs (178 / 168 / 126 / 96 registers for the residual):
// occupancy wins over spills up to 5 CTAs for the residual and 4 for the action (the default heuristic gave 126 / 142 registers).
#ifndef MADB_SF3D_MINB_RES
#define MADB_SF3D_MINB_RES 5
#endif
#ifndef MADB_SF3D_MINB_ACT
#define MADB_SF3D_MINB_ACT 4
#endif
template <int MODE> constexpr int sf3_minb() { return (MODE & MODE_ACT) ? MADB_SF3D_MINB_ACT : MADB_SF3D_MINB_RES; }
template <class Func, int ND, int NQ, int NEB, int MODE>
__global__ void __launch_bounds__(NQ *NQ *NEB, sf3_minb<MODE>()) k_sumfac3d(const __grid_constant__ Sf3Args<Func, ND, NQ> a)
{
   constexpr int ND3 = ND * ND * ND;
   const int tx = threadIdx.x, ty = threadIdx.y, tz = threadIdx.z;
   const int tid = ty * NQ + tx;
   const int t = a.begin + blockIdx.x * NEB + tz;
   const bool active = t < a.end;

   __shared__ double sU[NEB][ND3];          // dofs, [iz][iy][ix]
   __shared__ double sV[(MODE & MODE_ACT) ? NEB : 1][(MODE & MODE_ACT) ? ND3 : 1];
   __shared__ double sX[NEB][8][3];
   // scratch [iy|qy][qx][iz] with an ODD leading dimension in iz: consecutive threads (qx) are then LDZ doubles apart and hit
   // different banks (with LDZ = ND = 4 the 8-byte accesses of a half-warp fell on 4 banks: 38 M conflicts per colour launch)
   constexpr int LDZ = ND | 1;
   __shared__ double s0[NEB][NQ * NQ * LDZ];
   __shared__ double s1[NEB][NQ * NQ * LDZ];
   __shared__ double s2[NEB][NQ * NQ * LDZ];

   if (active)
   {
      for (int d = tid; d < ND3; d += NQ * NQ)
      {
         const int idx = a.vmap[(size_t)t * ND3 + d] & 0x7fffffff;
         sU[tz][d] = a.x[idx];
         if constexpr ((MODE & MODE_ACT) != 0) { sV[tz][d] = a.v[idx]; }
      }
      if (tid < 8)
      {
         const int n = a.e2n[(size_t)t * 8 + tid];
#pragma unroll
         for (int c = 0; c < 3; c++) { sX[tz][tid][c] = a.coords[(size_t)n * 3 + c]; }
      }
   }
   __syncthreads();

   // forward contractions of one dof array -> reference gradient at (tx,ty,qz), qz = 0..NQ-1
   auto forward = [&](const double *src, double (&gr)[NQ][3])
   {
      // x-step: thread (qx=tx, iy=ty<ND)
      if (ty < ND)
      {
#pragma unroll
         for (int iz = 0; iz < ND; iz++)
         {
            double sa = 0.0, sc = 0.0;
#pragma unroll
            for (int ix = 0; ix < ND; ix++)
            {
               const double uu = src[(iz * ND + ty) * ND + ix];
               sa = fma(a.B[tx][ix], uu, sa);
               sc = fma(a.G[tx][ix], uu, sc);
            }
            s0[tz][(ty * NQ + tx) * LDZ + iz] = sa;
            s1[tz][(ty * NQ + tx) * LDZ + iz] = sc;
         }
      }
      __syncthreads();
      // y-step: thread (qx=tx, qy=ty)
      double BB[ND], GB[ND], BG[ND];
#pragma unroll
      for (int iz = 0; iz < ND; iz++)
      {
         double bb = 0.0, gb = 0.0, bg = 0.0;
#pragma unroll
         for (int iy = 0; iy < ND; iy++)
         {
            const double av = s0[tz][(iy * NQ + tx) * LDZ + iz], cv = s1[tz][(iy * NQ + tx) * LDZ + iz];
            bb = fma(a.B[ty][iy], av, bb);
            gb = fma(a.B[ty][iy], cv, gb);
            bg = fma(a.G[ty][iy], av, bg);
         }
         BB[iz] = bb; GB[iz] = gb; BG[iz] = bg;
      }
      // z-step
#pragma unroll
      for (int qz = 0; qz < NQ; qz++)
      {
         double g0 = 0.0, g1 = 0.0, g2 = 0.0;
#pragma unroll
         for (int iz = 0; iz < ND; iz++)
         {
            g0 = fma(a.B[qz][iz], GB[iz], g0);
            g1 = fma(a.B[qz][iz], BG[iz], g1);
            g2 = fma(a.G[qz][iz], BB[iz], g2);
         }
         gr[qz][0] = g0; gr[qz][1] = g1; gr[qz][2] = g2;
      }
      __syncthreads();
   };

   double gu[NQ][3];
   forward(sU[tz], gu);
   double gv[(MODE & MODE_ACT) ? NQ : 1][3];
   if constexpr ((MODE & MODE_ACT) != 0) { forward(sV[tz], gv); }

   // trilinear geometry at (xi,eta) = (xq[tx], xq[ty]): bottom/top bilinear interpolants
   const double xi = a.xq[tx], eta = a.xq[ty];
   double a0[3], a1[3], b0[3], b1[3], dd[3];
#pragma unroll
   for (int c = 0; c < 3; c++)
   {
      const double v0 = sX[tz][0][c], v1 = sX[tz][1][c], v2 = sX[tz][2][c], v3 = sX[tz][3][c];
      const double v4 = sX[tz][4][c], v5 = sX[tz][5][c], v6 = sX[tz][6][c], v7 = sX[tz][7][c];
      a0[c] = (1.0 - eta) * (v1 - v0) + eta * (v3 - v2);
      a1[c] = (1.0 - eta) * (v5 - v4) + eta * (v7 - v6);
      b0[c] = (1.0 - xi) * (v2 - v0) + xi * (v3 - v1);
      b1[c] = (1.0 - xi) * (v6 - v4) + xi * (v7 - v5);
      const double p0 = (1.0 - eta) * ((1.0 - xi) * v0 + xi * v1) + eta * ((1.0 - xi) * v2 + xi * v3);
      const double p1 = (1.0 - eta) * ((1.0 - xi) * v4 + xi * v5) + eta * ((1.0 - xi) * v6 + xi * v7);
      dd[c] = p1 - p0;
   }

   Func f;
   f.load(a.fparams);
   double Z0[ND], Z1[ND], Z2[ND];
#pragma unroll
   for (int iz = 0; iz < ND; iz++) { Z0[iz] = 0.0; Z1[iz] = 0.0; Z2[iz] = 0.0; }
   double energy = 0.0;

#pragma unroll
   for (int qz = 0; qz < NQ; qz++)
   {
      const double zeta = a.xq[qz];
      double J[3][3], Ji[3][3], detJ;
#pragma unroll
      for (int c = 0; c < 3; c++)
      {
         J[c][0] = (1.0 - zeta) * a0[c] + zeta * a1[c];
         J[c][1] = (1.0 - zeta) * b0[c] + zeta * b1[c];
         J[c][2] = dd[c];
      }
      invert<3>(J, Ji, detJ);
      const double w = a.wq[tx] * a.wq[ty] * a.wq[qz] * detJ;
      double xin[3];
#pragma unroll
      for (int j = 0; j < 3; j++) { xin[j] = Ji[0][j] * gu[qz][0] + Ji[1][j] * gu[qz][1] + Ji[2][j] * gu[qz][2]; }
      if constexpr (MODE == MODE_ENERGY) { energy += f(xin, (const double *)nullptr) * w; }
      else
      {
         constexpr int ORDER = (MODE & MODE_ACT) ? 2 : 1;
         using T = AD<3, ORDER>;
         T xs[3];
#pragma unroll
         for (int m = 0; m < 3; m++) { xs[m] = ad_seed<3, ORDER>(xin[m], m); }
         const T res = f(xs, (const double *)nullptr);
         double gh[3];
         if constexpr ((MODE & MODE_ACT) != 0)
         {
            // y^ = w Ji H Ji^T v^ : physical direction, Hessian action, pull back
            double vp[3], hv[3];
#pragma unroll
            for (int j = 0; j < 3; j++) { vp[j] = Ji[0][j] * gv[qz][0] + Ji[1][j] * gv[qz][1] + Ji[2][j] * gv[qz][2]; }
#pragma unroll
            for (int i = 0; i < 3; i++)
            {
               ZD s {0.0, true};
#pragma unroll
               for (int j = 0; j < 3; j++) { s = zfmac(res.H(symidx_h<3>(i, j)), vp[j], s); }
               hv[i] = s.v;
            }
#pragma unroll
            for (int k = 0; k < 3; k++) { gh[k] = w * (Ji[k][0] * hv[0] + Ji[k][1] * hv[1] + Ji[k][2] * hv[2]); }
         }
         else
         {
#pragma unroll
            for (int k = 0; k < 3; k++) { gh[k] = w * (Ji[k][0] * res.g[0] + Ji[k][1] * res.g[1] + Ji[k][2] * res.g[2]); }
         }
#pragma unroll
         for (int iz = 0; iz < ND; iz++)
         {
            Z0[iz] = fma(a.B[qz][iz], gh[0], Z0[iz]);
            Z1[iz] = fma(a.B[qz][iz], gh[1], Z1[iz]);
            Z2[iz] = fma(a.G[qz][iz], gh[2], Z2[iz]);
         }
      }
   }

   if constexpr (MODE == MODE_ENERGY)
   {
      s0[tz][tid] = energy;
      __syncthreads();
      if (tid == 0 && active)
      {
         double e = 0.0;
         for (int k = 0; k < NQ * NQ; k++) { e += s0[tz][k]; }
         a.energy[t] = e;
      }
      return;
   }
   else
   {
      // reverse y-step: Z*(qx,qy,iz) -> Y0,Y1(qx,iy,iz)
#pragma unroll
      for (int iz = 0; iz < ND; iz++)
      {
         s0[tz][(ty * NQ + tx) * LDZ + iz] = Z0[iz];
         s1[tz][(ty * NQ + tx) * LDZ + iz] = Z1[iz];
         s2[tz][(ty * NQ + tx) * LDZ + iz] = Z2[iz];
      }
      __syncthreads();
      double Y0[ND], Y1[ND];
      if (ty < ND)
      {
#pragma unroll
         for (int iz = 0; iz < ND; iz++)
         {
            double y0 = 0.0, y1 = 0.0;
#pragma unroll
            for (int qy = 0; qy < NQ; qy++)
            {
               y0 = fma(a.B[qy][ty], s0[tz][(qy * NQ + tx) * LDZ + iz], y0);
               y1 = fma(a.G[qy][ty], s1[tz][(qy * NQ + tx) * LDZ + iz], y1);
               y1 = fma(a.B[qy][ty], s2[tz][(qy * NQ + tx) * LDZ + iz], y1);
            }
            Y0[iz] = y0; Y1[iz] = y1;
         }
      }
      __syncthreads();
      if (ty < ND)
      {
#pragma unroll
         for (int iz = 0; iz < ND; iz++)
         {
            s0[tz][(ty * NQ + tx) * LDZ + iz] = Y0[iz];
            s1[tz][(ty * NQ + tx) * LDZ + iz] = Y1[iz];
         }
      }
      __syncthreads();
      // reverse x-step + scatter: thread (ix=tx<ND, iy=ty<ND)
      if (active && tx < ND && ty < ND)
      {
#pragma unroll
         for (int iz = 0; iz < ND; iz++)
         {
            double out = 0.0;
#pragma unroll
            for (int qx = 0; qx < NQ; qx++)
            {
               out = fma(a.G[qx][tx], s0[tz][(ty * NQ + qx) * LDZ + iz], out);
               out = fma(a.B[qx][tx], s1[tz][(ty * NQ + qx) * LDZ + iz], out);
            }
            const int m = a.vmap[(size_t)t * ND3 + (iz * ND + ty) * ND + tx];
            const int idx = m & 0x7fffffff;
            a.y[idx] = (m < 0) ? out : a.y[idx] + out;
         }
      }
   }
}

template <class Func, int ND, int NQ> int launch_sumfac3d(const LaunchCtx &L, int mode)
{
#ifndef MADB_SF3D_CTA_THREADS
#define MADB_SF3D_CTA_THREADS 128 // target threads per CTA: NEB = that / NQ^2 elements per CTA
#endif
   constexpr int NEB = (MADB_SF3D_CTA_THREADS / (NQ * NQ) > 0) ? MADB_SF3D_CTA_THREADS / (NQ * NQ) : 1;
   static thread_local Sf3Args<Func, ND, NQ> a;
   if (mode & MODE_JAC) { return -2; }
   a.e2n = L.e2n; a.coords = L.coords; a.vmap = L.vmap;
   a.x = L.x; a.v = L.v; a.y = L.y; a.energy = L.energy;
   for (int i = 0; i < Func::N_PARAM; i++) { a.fparams[i] = L.fparams[i]; }
   for (int q = 0; q < NQ; q++)
   {
      for (int i = 0; i < ND; i++) { a.B[q][i] = L.b1d[0][q * ND + i]; a.G[q][i] = L.g1d[0][q * ND + i]; }
      a.xq[q] = L.xq1d[q];
      a.wq[q] = L.w1d[q];
   }
   const dim3 block(NQ, NQ, NEB);
   const int nlaunch = (mode == MODE_ENERGY) ? 1 : L.ncolors;
   for (int c = 0; c < nlaunch; c++)
   {
      a.begin = (mode == MODE_ENERGY) ? 0 : L.color_off[c];
      a.end = (mode == MODE_ENERGY) ? L.ne : L.color_off[c + 1];
      const int n = a.end - a.begin;
      if (n <= 0) { continue; }
      const int grid = (n + NEB - 1) / NEB;
      switch (mode)
      {
         case MODE_RES: k_sumfac3d<Func, ND, NQ, NEB, MODE_RES><<<grid, block, 0, L.stream>>>(a); break;
         case MODE_ACT: k_sumfac3d<Func, ND, NQ, NEB, MODE_ACT><<<grid, block, 0, L.stream>>>(a); break;
         case MODE_ENERGY: k_sumfac3d<Func, ND, NQ, NEB, MODE_ENERGY><<<grid, block, 0, L.stream>>>(a); break;
         default: return -1;
      }
   }
   return (int)cudaGetLastError();
}

template <class Func, int ND, int NQ> KernelOps make_ops_sumfac3d()
{
   KernelOps o;
   o.launch = &launch_sumfac3d<Func, ND, NQ>;
   o.n_input = 3;
   o.n_fparam = Func::N_PARAM;
   o.n_qprm = 0;
   o.n_field_qprm = 0;
   o.nvd = ND * ND * ND;
   o.ndof_all = ND * ND * ND;
   o.nq = NQ * NQ * NQ;
   o.ntab = ND * ND * ND;
   o.dim = 3;
   o.map_aos = 1;
   o.matrix_free_only = 1;
   return o;
}

/// scalar H1 space of order ND-1 on hexes, ADEval::GRAD, NQ^3 Gauss points
#define MADB_INSTANCE_SUMFAC3D(KIND, FUNC, ND, NQ)                                                                   \
   static ::madb::Registrar MADB_CAT(madb_reg3_, __COUNTER__)(                                                        \
      std::string(KIND) + "|d3q" + std::to_string(NQ) + "|" + std::to_string(ND) + ".1.4.0",                         \
      ::madb::make_ops_sumfac3d<FUNC, ND, NQ>());

} // namespace madb
