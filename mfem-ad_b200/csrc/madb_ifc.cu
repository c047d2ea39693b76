// madb_ifc.cu -- the interface reduction of the CSR values as a launch of its own (madb_integrator_assemble_end): between
// madb_integrator_assemble_begin and _end the caller starts the shared-dof exchange of the residual, whose NCCL transfer
// then runs under this kernel.
#include "madb_patch.cuh"

namespace madb
{
int launch_ifc_v(const PatchDev &P, double *vals, cudaStream_t stream)
{
   IfcList ly = P.ylist, lv = P.vlist;
   ly.n4 = ly.ng = 0;
   lv.stage = P.vstage;
   lv.out = vals;
   const int nb2 = (lv.n4 + 256 * IFC_U - 1) / (256 * IFC_U), nb3 = nb2 + (lv.ng + 255) / 256;
   if (nb3 > 0) { launch_pdl(k_ifc_reduce, nb3, 256, 0, stream, ly, lv, 0, 0, nb2); }
   return (int)cudaGetLastError();
}
} // namespace madb
