// madb_host.hpp -- host-side objects behind the C ABI (include/mfemad_b200.h).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace madb
{

enum { MODE_RES = 1, MODE_JAC = 2, MODE_ACT = 4, MODE_ENERGY = 8, MODE_COEF = 16 };

// ---- kernel registry -------------------------------------------------------------
// Fused kernels are templates on <functional type, element configuration>.  Each
// instantiation registers a launcher under a string key built from the same
// run-time descriptors the C ABI receives, e.g.
//    "minsurf|d2q4|3.1.4.0"          (kind | DIM,NQ1D | nd1d.vdim.mode.role per field)
// A missing key is a loud error naming the MADB_INSTANCE line to add.
// ---- patch assembly (madb_patch.cpp / madb_patch.cuh) ------------------------------
// Elements are grouped into compact patches of PATCH_PE elements = one CTA.  The threads of a
// CTA stage their element vectors / matrices in shared memory; every row of the patch is then
// summed from its sources in a fixed order and written once, coalesced, to y / the CSR values.
// Rows shared with other patches ("interface") go to a staging buffer and are summed in
// ascending patch order by a second small kernel.
constexpr int PATCH_PE = 128;          // elements per patch (small element matrices, one thread per element)
constexpr int PATCH_LD = PATCH_PE + 1; // leading dimension of the staged element data (bank spread)
constexpr int PATCH_PE_LARGE = 64;     // element matrices with more than 10 dofs: smaller patches, 4 threads per element
/// elements per patch / staging leading dimension for an element matrix of nvd dofs
constexpr int patch_pe(int nvd) { return nvd <= 10 ? PATCH_PE : PATCH_PE_LARGE; }
constexpr int patch_ld(int nvd) { return patch_pe(nvd) + 1; }
constexpr int PATCH_MAXEXTRA = 7;      // a slot has 1 + at most 7 sources (3-bit count)
inline constexpr int patch_al16(int bytes) { return (bytes + 15) & ~15; }
// the staged element matrices of one patch must fit in shared memory, two CTAs per SM
constexpr bool patch_eligible(int nvd) { return nvd <= 20; }
struct PatchDesc
{
   int ne;                    // elements of this patch (all PATCH_PE except possibly the last patch)
   int nrow_int, nrows;       // local rows: interior first, then interface
   int nyfold;                // words of the row fold list (8 phase counts + folds)
   int ystage_off;            // ystage[ystage_off + (lr - nrow_int)] <- interface rows
   int yblob_off, yblob_bytes; // residual-side maps of the patch: yblob + 16*yblob_off
   int nint, nexc, nslots;    // matrix slots: [0,nint) interior rows in CSR order, [nint,nexc) interface entries of
                              // this patch alone, [nexc,nslots) interface entries shared with other patches
   int nvfold;                // words of the slot fold list
   int nruns;                 // runs of consecutive CSR positions covering [0,nint)
   int stage_off;             // vstage[stage_off + (s - nexc)] <- shared interface slots
   int vblob_off, vblob_bytes; // matrix-side maps: vblob + 16*vblob_off
   int nchunk, nirr;          // chunk descriptors (padded to a multiple of 32) / irregular chunks (multiple of 8)
   int nvsrc;                 // entries of vsrc (>= nslots, padded to 32*nchunk)
   int npair, ngen;           // work lists of the write-out: aligned pairs of chunks / general chunks (each + 1 dummy entry)
};
// Blob layouts (sections padded to 16 bytes, copied to shared memory with one bulk copy each):
//   y blob: ysrc u16[nrows]  | ylist i32[nrow_int] | yfold u32[nyfold]
//   v blob: vsrc u16[nvsrc] | plist i32[2*(npair+1)] | glist i32[4*(ngen+1)] | isrc u16[32*nirr] | over i32[32*nirr] | vfold u32[nvfold]
//           (gather part first, fold list last: k_patch_ws fetches the fold list of the next patch as soon as the fold of
//           the current one is done and the gather part when its drain is complete; sizes: patch_yg_bytes / patch_vg_bytes)
//           the directly written slots [0,nexc) are cut into chunks of 32 slots; a chunk is written through one of
//           plist {g0, s0}: slots [s0, s0+64) (two full chunks) go to the 64 consecutive CSR positions from g0 (even):
//                 lane l gathers the slots s0 + 2l, s0 + 2l + 1 and writes them with one 16-byte store; g0 = -1: nothing
//           glist {g0, g1 - split, split, n | s0 << 8}: lanes < n store slot s0 + lane; lane < split -> CSR position
//                 g0 + lane, else g1 + (lane - split); n = 0: nothing
//           isrc / over: irregular chunks (more than one break or a dummy slot): sources / explicit positions, -1 = none
//           (both lists end with one entry that stores nothing: the device clamps list indices instead of branching)
// ysrc/vsrc: shared-memory location (entry * PATCH_LD + local element) of the first source of a row / slot.
// fold lists: 8 counts (phases 1..8), then words (dst | src << 16): staged[dst] += staged[src], phase by phase;
// phase k adds the k-th further source, so every row / slot is summed in ascending element order.
inline constexpr int patch_yg_bytes(const PatchDesc &D) { return patch_al16(2 * D.nrows) + patch_al16(4 * D.nrow_int); }
inline constexpr int patch_vg_bytes(const PatchDesc &D)
{
   return patch_al16(2 * D.nvsrc) + patch_al16(8 * (D.npair + 1)) + patch_al16(16 * (D.ngen + 1)) + patch_al16(64 * D.nirr) +
          patch_al16(128 * D.nirr);
}
struct IfcListDev // interface reduction lists of one side (see k_ifc_reduce)
{
   int n4, ng;
   const int4 *src4;
   const int *dst4;
   const int *ptr, *src, *dst;
   const double *stage;
   double *out;
};
// ---- CSR-image patch kernel (k_patch_img, madb_patch_img.cuh) -------------------------------------------------
// The threads of a patch scatter their element-matrix entries straight into a shared-memory IMAGE of the patch's part
// of the CSR value array (first source of a slot -> the slot; further sources -> private "extras" slots, folded onto
// the slot in ascending element order afterwards); the image is then written with bulk copies (cp.async.bulk
// shared -> global), one per run of consecutive CSR positions: no per-entry instructions on the write-out path.
// Image layout (doubles): [interior rows in CSR order, runs padded so that image and CSR positions have the same
// parity (16-byte alignment of the bulk copies)] [entries of interface rows only this patch contributes to]
// [entries shared with other patches: contiguous, even start and count -> one bulk copy to the staging buffer]
// [2 trash slots] [extras].  The y image: [local rows][1 trash][extras].
struct ImgDesc
{
   int ne, nrows, nrow_int, ystage_off;
   int mblob_off, mblob_bytes; // the patch's blob (one bulk copy to shared memory): mblob + 16*mblob_off
   int o_vmap, o_ymap, o_lists; // byte offsets of the scatter maps and of the lists inside the blob
   int nvfold, nyfold;         // fold lists: 8 phase counts + (dst | src << 16) words
   int nruns, nexcl;
   int sh0, nsh;               // shared interface entries: image range (even start, even count)
   int stage_off;              // vstage[stage_off + k] <- image[sh0 + k]
   int nvslots, nyslots;       // image sizes incl. trash and extras
   int pad[2];
};
// Blob of a patch: ImgDesc | vmap u32 [NEV][PE][TPE] (slot(i,j) | slot(j,i) << 16 of the entry thread (l, h) keeps at
// emission e) | ymap u16 [NEY][PE][TPE] | lists (ints, every section padded to a multiple of 4):
//   vfold [nvfold] | yfold [nyfold] | runs [4*nruns] {image offset, CSR position, count, 0} | excl_gpos [nexcl]
//   | excl_slot [u16 pairs, nexcl] | ylist [nrow_int]
struct ImgDev
{
   const ImgDesc *desc;
   const unsigned char *mblob;
   int max_vslots, max_yslots, max_mblob; // shared-memory sizing (slots / bytes)
   int nev, ney;                          // emissions per thread (matrix / vector)
};
constexpr int img_al4(int n) { return (n + 3) & ~3; }

struct PatchDev
{
   int npatch;
   int max_yblob, max_vblob; // bytes, shared-memory sizing
   int max_yg, max_yf, max_vg, max_vf; // the same split in gather part / fold list (k_patch_ws prefetches them separately)
   const PatchDesc *desc;
   const unsigned char *yblob, *vblob;
   double *ystage, *vstage;
   // interface reductions: out[dst] = sum of the staged partials, ascending patch order
   int ny_ifc, nv_ifc; // entries (statistics)
   IfcListDev ylist, vlist;
   ImgDev img; // CSR-image kernel (desc == null: not in use)
   int diag = 0; // MADB_DIAG (measurement only, results are wrong): 1 = writers skip the drain, 2 = compute warps skip the
                 // element computation
};

int launch_ifc_v(const PatchDev &P, double *vals, cudaStream_t stream); // madb_ifc.cu

struct LaunchCtx
{
   cudaStream_t stream;
   int ne, stride, ncolors;
   const int *color_off; // host, [ncolors+1] sorted-element offsets
   const int *e2n, *vmap, *pmap, *e2csr;
   const double *coords;
   const double *xe; // 2-D: vertex coordinates per element [4][stride][2] (null otherwise)
   const double *pdata[8];
   const double *qf;
   const double *x, *v;
   double *y, *vals, *energy;
   const int *perm;        // sorted position -> element (MODE_COEF output order)
   double *cvalue, *cgrad; // MODE_COEF outputs [e][q], [e][q][n]
   double *chess;          // MODE_COEF output [e][q][n][n] (null: not wanted)
   int coef_variant;       // 1: cgrad = ParamGradient::Eval as written (src/mmto.cpp:25-37)
   int write_y, write_vals;
   const double *fparams; // host
   // host tables, laid out exactly as madb::Tables<Cfg>
   const double *phi, *dphi, *gdphi, *w;
   // 1-D tables per field [nq1d][nd1d], 1-D points and weights (sum-factorised kernels)
   const double *b1d[8], *g1d[8];
   const double *xq1d, *w1d;
   const PatchDev *patch; // non-null: patch assembly (elements in patch order)
   int defer_v_ifc;       // 1: the interface reduction of the CSR values is left to launch_ifc_v (madb_integrator_assemble_end)
   cudaEvent_t ev0, ev1;  // non-null: recorded around the element kernel(s) (madb_integrator_set_timing)
};

constexpr int MADB_RC_MIRROR = -77; // KernelOps::launch: 1-D tables without mirror symmetry on the sum-factorised path

struct KernelOps
{
   int (*launch)(const LaunchCtx &, int mode);
   int n_input, n_fparam, n_qprm, n_field_qprm, nvd, ndof_all, nq, ntab, dim;
   int map_aos = 0;          // element maps stored [t][k] instead of [k][stride]
   int matrix_free_only = 0; // no assembled Jacobian (use grad_mult)
   int patch_ok = 0;         // patch-assembly kernels are compiled for this configuration
   int has_param_gradient = 0; // the functional implements ParamGradient::Eval as written (MODE_COEF variant 1)
   int patch_pe = 0;           // elements per patch of this configuration (0: patch_pe(nvd))
   int img_tpe = 0;            // CSR-image kernel: threads per element (0: kernel not compiled for this configuration)
   int img_mirror_nd = 0;      // img_tpe == 2: thread 1 works on the element mirrored in its second direction (1-D dofs)
};

std::map<std::string, KernelOps> &registry();
struct Registrar
{
   Registrar(const std::string &key, const KernelOps &ops);
};

// ---- 1-D bases and rules ---------------------------------------------------------
void gauss_legendre_01(int n, std::vector<double> &x, std::vector<double> &w);
void gauss_lobatto_01(int n, std::vector<double> &x);
void lagrange_tables(const std::vector<double> &nodes, const std::vector<double> &pts,
                     std::vector<double> &B, std::vector<double> &G); // [npts][nnodes]
inline int rule_npts_1d(int order) { return (order | 1) / 2 + 1; } // IntRules.Get(...): SURVEY a18

// ---- objects ---------------------------------------------------------------------
struct Ctx
{
   int device = 0;
   cudaStream_t stream = nullptr;
   double *scratch = nullptr, *scratch_host = nullptr; // reduction partials (device) and result (pinned host)
};

struct Mesh
{
   Ctx *ctx;
   int dim, ne, geom_order, nnodes;
   bool simplex = false;       // triangles (madb_mesh_create_simplex): 3 vertices per element, affine map
   std::vector<int> e2n;       // [ne][ngn()]: 2^dim vertices lexicographic, or the dim+1 vertices of a simplex
   std::vector<double> coords; // [nnodes][dim]
   double *d_coords = nullptr;
   int ngn() const { return simplex ? dim + 1 : (1 << dim); }
};

// ---- triangles: rules and nodal bases as MFEM defines them (restated; MFEM is not available here) ------------------------
/// IntRules.Get(Geometry::TRIANGLE, order), orders 0 - 6: points (x, y) on the reference triangle (0,0),(1,0),(0,1),
/// weights summing to 1/2.  Returns false for orders without a table here.
bool triangle_rule(int order, std::vector<double> &pts, std::vector<double> &w);
/// number of scalar dofs of the order-p nodal space on a triangle: H1 p = 1, 2 (vertices, then edge midpoints in edge
/// order (0,1), (1,2), (2,0)); L2 p = 0 (constant).  0: not available.
int triangle_ndof(int basis, int order);
/// shape functions and reference derivatives at (x, y): phi[nd], dphi[nd][2]
void triangle_shapes(int basis, int order, double x, double y, double *phi, double *dphi);

enum { BASIS_H1 = 0, BASIS_L2 = 1 };
enum { ORD_BYNODES = 0, ORD_BYVDIM = 1 };

struct Space
{
   Ctx *ctx;
   Mesh *mesh;
   int basis, order, vdim, ordering;
   int ndofs;            // scalar dofs
   std::vector<int> e2l; // [ne][(order+1)^dim] lexicographic scalar dof ids
   int nd_el() const
   {
      if (mesh->simplex) { return triangle_ndof(basis, order); }
      int n = 1;
      for (int d = 0; d < mesh->dim; d++) { n *= (order + 1); }
      return n;
   }
   int vsize() const { return ndofs * vdim; }
};

struct Functional
{
   std::string kind;           // e.g. "minsurf", "pg"
   std::vector<double> params; // own constants
   std::vector<int> iparams;   // structural integers (become part of the key)
   std::vector<Functional *> children;
   std::string key() const;
   void flat_params(std::vector<double> &out) const;
};

struct FieldDesc
{
   Space *space;
   unsigned mode;
   int role; // ROLE_INPUT / ROLE_PARAM
   bool qvalue = false; // declared as ADEval::QVALUE (treated as VALUE on the rule's nodal L2 space)
};

struct Integrator
{
   Ctx *ctx = nullptr;
   Mesh *mesh = nullptr;
   std::vector<FieldDesc> fields;
   Functional *fn = nullptr;
   int quad_order = -1, nq1d = 0, nq = 0;
   std::string key;
   KernelOps ops;

   // sizes
   int nvd = 0, ndof_all = 0, npd = 0;
   long ntotal = 0;        // size of the concatenated vector
   std::vector<long> goff; // block offset of each input field
   int ne = 0, stride = 0;

   // colouring / ordering
   std::vector<int> perm;      // sorted position -> element
   std::vector<int> color_off; // [ncolors+1]

   // pattern (host) -- built lazily
   bool have_pattern = false;
   std::vector<int> rowptr, colidx;

   // device data
   int *d_e2n = nullptr, *d_vmap = nullptr, *d_pmap = nullptr, *d_e2csr = nullptr;
   double *d_xe = nullptr; // 2-D: vertex coordinates per element, [4][stride][2]
   long untouched_rows = 0; // dofs of the concatenated vector that belong to no element (residual zeroed before assembly then)
   double *pending_vals = nullptr; // madb_integrator_assemble_begin: CSR values whose interface reduction is still to be launched
   int *d_rowptr = nullptr, *d_colidx = nullptr, *d_perm = nullptr;
   double *d_cvalue = nullptr, *d_cgrad = nullptr, *d_chess = nullptr;
   double *d_energy = nullptr, *d_esum = nullptr;
   double *d_x = nullptr, *d_v = nullptr, *d_v2 = nullptr, *d_y = nullptr, *d_vals = nullptr; // staging for host callers
   std::vector<double *> d_pstage;                                           // staging of parameter fields
   double *d_qf = nullptr;
   size_t qf_count = 0;
   std::vector<const double *> pdata; // device pointers of the parameter fields (per field)
   std::vector<double> phi, dphi, gdphi, w;
   std::vector<std::vector<double>> b1d, g1d;
   std::vector<double> xq1d, w1d;
   std::vector<double> tri_pts; // simplex meshes: reference points of the rule [nq][2]

   // patch assembly
   bool use_patches = false;
   int pe = PATCH_PE;           // elements per patch (patch_pe(nvd))
   std::vector<PatchDesc> pdesc;
   std::vector<int> prows;      // concatenated local row lists (global dof ids), per patch [nrows]
   std::vector<int> prow_off;   // [npatch+1]
   int max_yblob = 0, max_vblob = 0;
   int max_yg = 0, max_yf = 0, max_vg = 0, max_vf = 0;
   bool have_patch_vals = false;
   PatchDev pdev {};
   PatchDesc *d_pdesc = nullptr;
   unsigned char *d_yblob = nullptr, *d_vblob = nullptr;
   ImgDesc *d_idesc = nullptr; // CSR-image kernel maps
   unsigned char *d_mblob = nullptr;
   long img_map_bytes = 0; // statistics: scatter maps + lists (bytes read per assembly)
   double *d_ystage = nullptr, *d_vstage = nullptr;
   int *d_ifc[2][5] = {{nullptr, nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr, nullptr}}; // [side][src4, dst4, ptr, src, dst]

   // optional device timing of the element kernel(s)
   bool timing = false;
   cudaEvent_t ev0 = nullptr, ev1 = nullptr;

   // essential dofs
   int ness = 0;
   int *d_ess = nullptr;

   ~Integrator();
};

const char *last_error();
void set_error(const std::string &s);

// pattern / colouring helpers (madb_pattern.cpp)
void build_vdofs(const Integrator &I, int e, std::vector<int> &vd);
void color_elements(const Integrator &I, std::vector<int> &color, int &ncolors);
void build_pattern(Integrator &I);
void build_e2csr(const Integrator &I, const std::vector<int> &color, std::vector<int> &e2csr);

// patch assembly (madb_patch.cpp)
struct PatchHost // maps of one side (residual or matrix)
{
   std::vector<unsigned char> blob;
   std::vector<int> ptr, src, dst; // interface reduction lists: entries with more than 4 sources
   std::vector<int> src4, dst4;    // packed entries: 4 sources (-1: none) each
   long stage_size = 0;
};
void patch_order(Integrator &I);                      // fills I.perm (patch order) and I.pdesc[].ne
bool patch_build_y(Integrator &I, PatchHost &H);       // needs I.perm; false: not representable
bool patch_build_v(Integrator &I, PatchHost &H);       // needs the CSR pattern
struct ImgHost // maps of the CSR-image kernel
{
   std::vector<ImgDesc> desc;
   std::vector<unsigned char> mblob;
   int max_vslots = 0, max_yslots = 0, max_mblob = 0, nev = 0, ney = 0;
};
/// emission schedule of the element threads: local (I, J) of kept matrix entry e, local I of kept vector entry e
void img_schedule(int nvd, int tpe, int mirror_nd, std::vector<int> &vI, std::vector<int> &vJ, std::vector<int> &yI);
bool patch_build_img(Integrator &I, PatchHost &H, ImgHost &IH); // needs the CSR pattern and patch_build_y
int patch_selftest(Mesh &mesh, Space &space, double *max_err, long *stats); // host-only emulation (CPU tests)
int patch_selftest_img(Mesh &mesh, Space &space, int tpe, double *max_err, long *stats); // same for the CSR-image kernel

} // namespace madb
