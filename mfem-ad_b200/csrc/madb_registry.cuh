// madb_registry.cuh -- glue between the typed kernels and the run-time registry.
#pragma once
#include "madb_host.hpp"
#include "madb_kernels.cuh"
#include "madb_patch.cuh"
#include "madb_patch_img.cuh"

#include <cstring>
#include <string>

namespace madb
{

template <class Cfg> std::string config_key()
{
   // tensor elements: "d<dim>q<points per direction>|<dofs per direction>...."; simplices: "s<dim>q<points>|<dofs per element>...."
   std::string k = Cfg::TENSOR ? "d" + std::to_string(Cfg::DIM) + "q" + std::to_string(Cfg::NQ1D)
                               : "s" + std::to_string(Cfg::DIM) + "q" + std::to_string(Cfg::NQ);
   static_for<Cfg::NF>([&](auto F)
   {
      constexpr int fi = decltype(F)::value;
      using Fd = typename Cfg::template field<fi>;
      k += "|" + std::to_string(Cfg::TENSOR ? Fd::ND1D : Cfg::template nd<fi>()) + "." + std::to_string(Fd::VDIM) + "." +
           std::to_string((int)(Fd::MODE & (EV_VALUE | EV_GRAD))) + "." + std::to_string(Fd::ROLE);
   });
   return k;
}

template <class Func, class Cfg> AsmArgs<Func, Cfg> &fill_args(const LaunchCtx &L)
{
   using Args = AsmArgs<Func, Cfg>;
   static thread_local Args a; // ~tens of KB: keep off the stack; one block per host thread (one thread drives a context)
   a.stride = L.stride;
   a.write_y = L.write_y;
   a.write_vals = L.write_vals;
   a.e2n = L.e2n;
   a.coords = L.coords;
   a.xe = L.xe;
   a.vmap = L.vmap;
   a.pmap = L.pmap;
   for (int f = 0; f < Cfg::NF; f++) { a.pdata[f] = L.pdata[f]; }
   a.qf = L.qf;
   a.e2csr = L.e2csr;
   a.x = L.x;
   a.v = L.v;
   a.y = L.y;
   a.vals = L.vals;
   a.energy = L.energy;
   a.perm = L.perm;
   a.cvalue = L.cvalue;
   a.cgrad = L.cgrad;
   a.chess = L.chess;
   a.coef_variant = L.coef_variant;
   for (int i = 0; i < Func::N_PARAM; i++) { a.fparams[i] = L.fparams[i]; }
   std::memcpy(a.tab.phi, L.phi, sizeof(a.tab.phi));
   std::memcpy(a.tab.dphi, L.dphi, sizeof(a.tab.dphi));
   std::memcpy(a.tab.gdphi, L.gdphi, sizeof(a.tab.gdphi));
   std::memcpy(a.tab.w, L.w, sizeof(a.tab.w));
   fill_sf2d(a.sf, L.b1d[0], L.g1d[0], L.xq1d, L.w1d);
   return a;
}

template <class Func, class Cfg, bool UNROLLQ> int launch_impl(const LaunchCtx &L, int mode)
{
   using Args = AsmArgs<Func, Cfg>;
   Args &a = fill_args<Func, Cfg>(L);
   // the even / odd contraction of the sum-factorised 2-D path needs mirror-symmetric 1-D nodes and points (every
   // tensor basis and Gauss rule of MFEM is): refuse anything else loudly instead of assembling wrong values
   if (!sf2d_mirror_ok(a.sf)) { return MADB_RC_MIRROR; }
   if constexpr (patch_eligible(Cfg::NVD))
   {
      if (L.patch && mode != MODE_ENERGY && mode != MODE_COEF)
      {
         a.begin = 0;
         a.end = L.ne;
         switch (mode)
         {
            case MODE_RES: return launch_patch_mode<Func, Cfg, MODE_RES, UNROLLQ>(a, L);
            case MODE_RES | MODE_JAC: return launch_patch_mode<Func, Cfg, MODE_RES | MODE_JAC, UNROLLQ>(a, L);
            case MODE_ACT: return launch_patch_mode<Func, Cfg, MODE_ACT, UNROLLQ>(a, L);
            default: return -1;
         }
      }
   }
   const bool whole = (mode == MODE_ENERGY || mode == MODE_COEF);
   const int nlaunch = whole ? 1 : L.ncolors;
   if (L.ev0) { cudaEventRecord(L.ev0, L.stream); }
   for (int c = 0; c < nlaunch; c++)
   {
      a.begin = whole ? 0 : L.color_off[c];
      a.end = whole ? L.ne : L.color_off[c + 1];
      const int n = a.end - a.begin;
      if (n <= 0) { continue; }
      const int grid = (n + 127) / 128;
      switch (mode)
      {
         case MODE_RES: k_element<Func, Cfg, MODE_RES, UNROLLQ><<<grid, 128, 0, L.stream>>>(a); break;
         case MODE_RES | MODE_JAC:
            k_element<Func, Cfg, MODE_RES | MODE_JAC, UNROLLQ><<<dim3(grid, element_parts<Cfg, MODE_RES | MODE_JAC>()), 128, 0, L.stream>>>(a);
            break;
         case MODE_ACT: k_element<Func, Cfg, MODE_ACT, UNROLLQ><<<grid, 128, 0, L.stream>>>(a); break;
         case MODE_ENERGY: k_element<Func, Cfg, MODE_ENERGY, UNROLLQ><<<grid, 128, 0, L.stream>>>(a); break;
         case MODE_COEF: k_element<Func, Cfg, MODE_COEF, UNROLLQ><<<grid, 128, 0, L.stream>>>(a); break;
         default: return -1;
      }
   }
   if (L.ev1) { cudaEventRecord(L.ev1, L.stream); }
   return (int)cudaGetLastError();
}

template <class Func, class Cfg, bool UNROLLQ> KernelOps make_ops()
{
   KernelOps o;
   o.map_aos = 0;
   o.matrix_free_only = 0;
   o.patch_ok = patch_eligible(Cfg::NVD) ? 1 : 0;
   o.patch_pe = patch_pe_of<Func, Cfg>();
   o.img_tpe = img_eligible<Func, Cfg>() ? img_tpe<Func, Cfg>() : 0;
   o.img_mirror_nd = (o.img_tpe == 2) ? Cfg::template field<0>::ND1D : 0;
   o.has_param_gradient = has_param_gradient<Func>::value ? 1 : 0;
   o.launch = &launch_impl<Func, Cfg, UNROLLQ>;
   o.n_input = Cfg::N_INPUT;
   o.n_fparam = Func::N_PARAM;
   o.n_qprm = Func::N_QPRM;
   o.n_field_qprm = Cfg::N_FIELD_QPRM;
   o.nvd = Cfg::NVD;
   o.ndof_all = Cfg::NDOF_ALL;
   o.nq = Cfg::NQ;
   o.ntab = Cfg::NTAB;
   o.dim = Cfg::DIM;
   return o;
}

#define MADB_CAT2(a, b) a##b
#define MADB_CAT(a, b) MADB_CAT2(a, b)
/// Register the fused kernels of functional type FUNC (run-time key KIND) on
/// element configuration CFG (a madb::Config<...>; wrap in parentheses-free alias).
#define MADB_INSTANCE(KIND, FUNC, CFG, UNROLLQ)                                                         \
   static ::madb::Registrar MADB_CAT(madb_reg_, __COUNTER__)(std::string(KIND) + "|" + ::madb::config_key<CFG>(), \
                                                             ::madb::make_ops<FUNC, CFG, UNROLLQ>());

/// Single vector space with ADEval::VECTOR: the variant that reproduces the reference's single-space
/// AssembleElementGrad arithmetic (src/ad_intg.hpp:310-326, RefVectorOf in madb_functionals.cuh), registered under the
/// same kind with the key suffix "|refvec".  madb_integrator_create selects it for one-input-space VECTOR integrators
/// unless MADB_INTEG_BLOCK asks for the index-consistent (ADBlockNonlinearFormIntegrator) contraction.
#define MADB_INSTANCE_REFVEC(KIND, FUNC, CFG, UNROLLQ)                                                             \
   static ::madb::Registrar MADB_CAT(madb_regrv_, __COUNTER__)(                                                    \
      std::string(KIND) + "|" + ::madb::config_key<CFG>() + "|refvec",                                             \
      ::madb::make_ops<::madb::RefVectorOf<FUNC, CFG::template sd<0>(), CFG::template field<0>::VDIM>, CFG, UNROLLQ>());

} // namespace madb
