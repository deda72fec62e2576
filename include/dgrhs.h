/*
 * dgrhs.h -- C-ABI of the B200-native DG evolution right-hand side.
 *
 * This is the drop-in boundary for ONE path of SpECTRE (reference checkout
 * v2024.09.29, paths relative to its root): per-element volume time derivative
 * + boundary corrections + lift + time-stepper substep for the ScalarWave and
 * GeneralizedHarmonic systems on Mesh<3>.
 *
 * Conventions
 *  - every entry point returns 0 on success; on failure it returns non-zero and
 *    dgrhs_last_error() holds a message.  The reference has no error codes on
 *    this path (ASSERT/ERROR abort, Utilities/ErrorHandling/Error.hpp:68-77);
 *    the C++ shims in spectre_b200/host/ turn non-zero into an exception.
 *  - plain pointers and sizes only.  "host" pointers are ordinary host memory
 *    in the reference's Variables layout (DataStructures/Variables.hpp:94-160):
 *    per element one contiguous block, component-major, n = N^3 points per
 *    component, grid index i + N*(j + N*k).  Tensor component order follows
 *    Tensor/Structure.hpp:162-194 (see DESIGN.md "Data layout").
 *  - evolved variables per point: ScalarWave 5 (Psi, Pi, Phi_i), GH 50
 *    (g_ab 10, Pi_ab 10, Phi_iab 30).
 *  - the library owns all device memory; element data stays resident in HBM.
 *  - one context per GPU, driven from one host thread (the reference runs one
 *    element per Charm++ PE, single-threaded: SURVEY.md 8b "Threading").
 */
#ifndef DGRHS_H
#define DGRHS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dgrhs_ctx dgrhs_ctx;

enum { DGRHS_SYSTEM_SCALAR_WAVE = 0, DGRHS_SYSTEM_GH = 1 };

/* gh::gauges::GaugeCondition subclasses
 * (Evolution/Systems/GeneralizedHarmonic/GaugeSourceFunctions/) */
enum {
  DGRHS_GAUGE_HARMONIC = 0,        /* Harmonic.cpp:25-39 */
  DGRHS_GAUGE_FIELDS = 1,          /* H_a, d_a H_b supplied as per-point fields:
                                      AnalyticChristoffel.cpp:64-149 evaluated by
                                      the host (static solutions) */
  DGRHS_GAUGE_DAMPED_HARMONIC = 2, /* DampedHarmonic.cpp:70-439 */
  DGRHS_GAUGE_ANALYTIC_GAUGE_WAVE = 3 /* AnalyticChristoffel with the GaugeWave
                                      solution, re-evaluated on the device at the
                                      time of every RHS call */
};

/* Spectral::MortarSize (Spectral/SegmentSize.hpp) */
enum { DGRHS_MORTAR_FULL = 0, DGRHS_MORTAR_LOWER_HALF = 1, DGRHS_MORTAR_UPPER_HALF = 2 };

/* TimeSteppers (Time/TimeSteppers/) */
enum {
  DGRHS_STEPPER_ADAMS_BASHFORTH = 0, /* AdamsBashforth.cpp:120-201, order 1..8 */
  DGRHS_STEPPER_RK3_HESTHAVEN = 1,   /* Rk3HesthavenSsp.cpp:55-81 */
  /* RungeKutta::update_u_impl with a Butcher tableau (RungeKutta.cpp:69-122) */
  DGRHS_STEPPER_RK3_OWREN = 2,       /* Rk3Owren.cpp:17-34 */
  DGRHS_STEPPER_RK3_KENNEDY = 3,     /* Rk3Kennedy.cpp:18-43 (explicit part) */
  DGRHS_STEPPER_RK4 = 4,             /* ClassicalRungeKutta4.cpp:24-49 */
  DGRHS_STEPPER_DORMAND_PRINCE5 = 5  /* DormandPrince5.cpp:21-50 */
};

const char* dgrhs_last_error(void);

/* Number of CUDA kernels this library has launched in this process so far
 * (bench.py reports the difference over the timed region). */
int64_t dgrhs_kernel_launch_count(void);

/* ---- context ----------------------------------------------------------- */

/* Replaces the per-element DataBox of the reference's DgElementArray
 * (Evolution/DiscontinuousGalerkin/DgElementArray.hpp:68-100) by one batched
 * structure-of-arrays block per GPU.  n_points_1d = N of the isotropic
 * Legendre-Gauss-Lobatto Mesh<3> (ComputeTimeDerivative.hpp:415-430 asserts the
 * same restriction).  n_ghost_faces = number of mortar faces whose neighbour
 * lives on another rank (0 for single GPU). */
int dgrhs_create(dgrhs_ctx** ctx, int system, int n_points_1d, int n_elements,
                 int n_ghost_faces, int device);
int dgrhs_destroy(dgrhs_ctx* ctx);

/* Geometry consumed (not computed) by the path, SURVEY.md 8 a22:
 *  inv_jacobian: host [n_elements][9][n], component ihat + 3*i  (=
 *     domain::Tags::InverseJacobian<3, ElementLogical, Inertial>, TypeAliases
 *     .hpp:444-447)
 *  coords: host [n_elements][3][n] inertial coordinates, or NULL when no
 *     consumer needs them
 *  neighbors: host [n_elements][6], direction d = 2*dim + side (side 0 = lower):
 *     >= 0 local element index of the aligned, conforming neighbour
 *     (Domain/Structure/Element.hpp neighbours + OrientationMap::is_aligned);
 *     -1 external boundary (no correction); <= -2: ghost face -(value+2), whose
 *     neighbour data arrive through the halo buffers below. */
int dgrhs_set_geometry(dgrhs_ctx* ctx, const double* inv_jacobian,
                       const double* coords, const int32_t* neighbors);

/* Non-aligned neighbours (rotated blocks, e.g. the wedges of a Sphere): the
 * face-restricted OrientationMap of every neighbour (Domain/Structure/
 * OrientationMap.hpp, orient_variables_on_slice in OrientationMapHelpers.cpp:25-
 * 120).  neighbor_direction host [n_elements][6]: the neighbour's direction
 * (0..5) whose face touches ours; face_permutation host [n_elements][6]: how our
 * face point (qa, qb) -- the two remaining logical dimensions in increasing
 * order -- maps to the neighbour's: bit0 swap (qa, qb), bit1 / bit2 reverse the
 * neighbour's first / second face coordinate.  Evolved tensors have inertial
 * components, so only the point index is transformed.  Optional: without this
 * call neighbours are aligned (direction d^1, identity).  Same N on both sides. */
int dgrhs_set_neighbor_orientations(dgrhs_ctx* ctx, const int32_t* neighbor_direction,
                                    const int32_t* face_permutation);

/* Static per-point fields: ScalarWave: gamma2 (1 component,
 * ScalarWave/Initialize.hpp:48-49); GH: gamma0, gamma1, gamma2 (3 components,
 * GeneralizedHarmonic/Initialize.hpp:59-71).  host [n_elements][ncomp][n]. */
int dgrhs_set_static_fields(dgrhs_ctx* ctx, const double* fields, int ncomp);

/* ---- Local time stepping ---------------------------------------------
 * TimeSteppers::adams_lts::lts_coefficients, explicit (Adams-Bashforth) schemes
 * (src/Time/TimeSteppers/AdamsLts.cpp:330-437): the nonzero terms of the boundary
 * contribution to the local side's step from start_tick to end_tick.  Times are integer ticks
 * of tick_size after time_origin, in the order of insertion into the boundary history
 * (ConstBoundaryHistoryTimes); term t couples local_ticks[local_index[t]] with
 * remote_ticks[remote_index[t]], sorted like the reference's LtsCoefficients.  Host only. */
int dgrhs_adams_lts_coefficients(int local_order, int remote_order, int small_step_order,
                                 int n_local, const long long* local_ticks, int n_remote,
                                 const long long* remote_ticks, long long start_tick,
                                 long long end_tick, double time_origin, double tick_size,
                                 int max_terms, int* n_terms, int* local_index,
                                 int* remote_index, double* coefficients);
/* The same with the scheme types of AdamsLts.hpp:84-89: *_implicit != 0 selects an
 * Adams-Moulton (implicit) scheme for that side / for the small steps.  An id is the step id
 * at ticks[i] when substep_sizes[i] == 0 (or substep_sizes == NULL), else the substep
 * (predictor) id of the step from ticks[i] to ticks[i] + substep_sizes[i], listed after its
 * step id like BoundaryHistory holds it. */
int dgrhs_adams_lts_coefficients_general(
    int local_implicit, int local_order, int remote_implicit, int remote_order,
    int small_step_implicit, int small_step_order, int n_local, const long long* local_ticks,
    const long long* local_substep_sizes, int n_remote, const long long* remote_ticks,
    const long long* remote_substep_sizes, long long start_tick, long long end_tick,
    double time_origin, double tick_size, int max_terms, int* n_terms, int* local_index,
    int* remote_index, double* coefficients);
/* Adams-Bashforth local time stepping with fixed step sizes dt_coarse / 2^levels[e]
 * (levels ascending in the element order: coarse steps first), replacing the action pair
 * UpdateU + ApplyLtsBoundaryCorrections (Actions/UpdateU.hpp:44-120,
 * ApplyBoundaryCorrections.hpp:1142-1192; AdamsBashforth::add_boundary_delta_impl,
 * AdamsBashforth.cpp:264-281): the volume part of the time derivative incl. the external
 * boundary conditions goes through each element's own history, the boundary corrections of
 * internal faces are integrated from the histories of both sides.  The step sizes do not
 * change (no step choosers), the histories start from given past states like
 * TimeStepperTestUtils::initialize_history: after dgrhs_set_state(u(t0)) call
 * dgrhs_lts_set_past_state for j = 1 .. order-1 with, for every element, its state at
 * t0 - j * (its own step).  One GPU; conforming faces (aligned or oriented) and non-conforming
 * 2:1 mortars (dgrhs_set_mortars); ghost (DirichletAnalytic) boundary conditions with static
 * data, ConstraintPreservingBjorhus and DemandOutgoingCharSpeeds faces; static gauge fields;
 * the exponential filter after each element's step; no p-mortars, no moving mesh.  Time runs in ticks of the
 * finest step; the state is complete (all elements at the same time) after a multiple of
 * dgrhs_lts_ticks_per_coarse_step ticks. */
int dgrhs_lts_init(dgrhs_ctx* ctx, int order, double t0, double dt_coarse,
                   const int32_t* levels);
/* Takes effect at the next dgrhs_lts_init.  0: every internal face goes through the boundary histories, the
 * reference's formulation term by term.  1 (default): faces between elements of the same
 * level are evaluated like GTS faces -- their corrections enter the volume history, which is
 * what lts_coefficients_for_gts (AdamsLts.cpp:307-327) sums to, in another order of additions
 * -- and the stepper update is fused into the volume kernel (orders <= 4); only faces between
 * different levels keep histories. */
int dgrhs_lts_set_mode(dgrhs_ctx* ctx, int same_level_faces_in_volume_history);
int dgrhs_lts_set_past_state(dgrhs_ctx* ctx, int j, const double* u_past);
int dgrhs_lts_take_ticks(dgrhs_ctx* ctx, long long n_ticks);
int dgrhs_lts_ticks_per_coarse_step(dgrhs_ctx* ctx, long long* n_ticks);
int dgrhs_lts_time(dgrhs_ctx* ctx, double* time, long long* tick);

/* Moving mesh: inertial mesh velocity v_g^i at the grid points, host [n_elements][3][n]
 * (Tags::MeshVelocity), or NULL for a static mesh (the default).  With a velocity set the
 * right-hand side gains the terms of a moving mesh for systems without fluxes: dt u += v_g^i
 * d_i u (VolumeTermsImpl.tpp:155-235), GH: gamma1 v_g.C3 in dt g and gamma1 gamma2 v_g.C3 in
 * dt Pi (GeneralizedHarmonic/TimeDerivative.cpp:237-300,372-378), and the characteristic
 * speeds of dg_package_data are taken relative to the mesh (normal_dot_mesh_velocity,
 * UpwindPenalty.cpp:85-91 / ScalarWave UpwindPenalty.cpp:55-67; also for
 * DemandOutgoingCharSpeeds).  The caller supplies the inverse Jacobian (dgrhs_set_geometry) and
 * the velocity of the current time; the maps themselves stay with the caller.  Conforming
 * faces and ghost boundary conditions only (Bjorhus faces and non-conforming mortars are
 * rejected); the stepper update is not fused on a moving mesh. */
int dgrhs_set_mesh_velocity(dgrhs_ctx* ctx, const double* mesh_velocity);

/* GH gauge condition.  params: DAMPED_HARMONIC: {width, amp_L1, amp_L2, amp_S,
 * exp_L1, exp_L2, exp_S}; ANALYTIC_GAUGE_WAVE: {amplitude, wavelength}. */
int dgrhs_set_gauge(dgrhs_ctx* ctx, int gauge, const double* params, int nparams);
/* DGRHS_GAUGE_FIELDS: gauge_h host [n_elements][4][n]; d4_gauge_h host
 * [n_elements][16][n] with d_a H_b at component a + 4*b (tnsr::ab). */
int dgrhs_set_gauge_fields(dgrhs_ctx* ctx, const double* gauge_h,
                           const double* d4_gauge_h);
/* AnalyticChristoffel gauge for a STATIC analytic solution (AnalyticChristoffel
 * .cpp:64-149): u_analytic host [n_elements][50][n] = the solution's (g, Pi,
 * Phi); H_a = -Gamma_a and its numerical spatial derivative are evaluated once
 * on the device and kept as gauge fields (d_t H_a = 0).  Needs set_geometry. */
int dgrhs_set_gauge_analytic_christoffel(dgrhs_ctx* ctx, const double* u_analytic);

/* Evolved variables, host [n_elements][n_vars][n] (Variables layout). */
int dgrhs_set_state(dgrhs_ctx* ctx, const double* u);
int dgrhs_get_state(dgrhs_ctx* ctx, double* u);
/* Stream-ordered forms of the two calls above: the copy is queued behind the
 * context's earlier work (steps included) and the call returns immediately;
 * `u` must stay valid, and for a copy that really overlaps other contexts'
 * work be page-locked, until dgrhs_synchronize(ctx) returns.  Two contexts
 * driven this way double-buffer independent batches: one uploads while the
 * other steps and downloads (PCIe is full duplex). */
int dgrhs_set_state_async(dgrhs_ctx* ctx, const double* u);
int dgrhs_get_state_async(dgrhs_ctx* ctx, double* u);
/* Last computed time derivative (after boundary corrections), same layout. */
int dgrhs_get_time_derivative(dgrhs_ctx* ctx, double* dt_u);

/* ---- the hot path ------------------------------------------------------ */

/* = evolution::dg::Actions::ComputeTimeDerivative (ComputeTimeDerivative.hpp:
 * 383-650: partial_derivatives + System::TimeDerivative::apply, face
 * projection, normals, dg_package_data) followed by
 * ApplyBoundaryCorrectionsToTimeDerivative (ApplyBoundaryCorrections.hpp:
 * 1093-1126: dg_boundary_terms, lift_flux, add_slice_to_data) for every element
 * of the batch.  volume_only != 0 skips the boundary corrections (unit tests). */
int dgrhs_compute_time_derivative(dgrhs_ctx* ctx, double time, int volume_only);

/* Multi-GPU split of the same call: elements [0, n_interior)
 * have no ghost faces.  Sequence per RHS: pack_halo -> exchange (caller, NCCL)
 * -> compute(interior) may overlap -> compute(boundary) after the exchange.
 * An empty range is a no-op; the first non-empty range of an evaluation counts it
 * (dgrhs_rhs_evaluations) and runs the once-per-RHS kernels. */
int dgrhs_set_interior_count(dgrhs_ctx* ctx, int n_interior);
int dgrhs_pack_halo(dgrhs_ctx* ctx);
int dgrhs_compute_time_derivative_range(dgrhs_ctx* ctx, double time,
                                        int elem_begin, int elem_end);
/* Device pointers of the halo buffers: send[n_ghost_faces][halo_comps][N^2]
 * and recv (same shape); ghost_send_map host [n_send][2] = {local element,
 * direction} whose face is packed into send slot i (n_send <= n_ghost_faces;
 * receive slots beyond the exchanged ones hold boundary-condition data).  Static neighbour-side
 * face data (inverse-Jacobian row, gamma1, gamma2) travel in the same slots, so
 * one exchange per RHS suffices. */
int dgrhs_set_halo_map(dgrhs_ctx* ctx, const int32_t* ghost_send_map, int n_send);
/* External boundaries with a ghost boundary condition (SURVEY 8f rank 1:
 * apply_boundary_conditions_on_all_external_faces, BoundaryConditionsImpl.hpp:
 * 672-764, with BoundaryCondition::dg_ghost such as GeneralizedHarmonic/
 * BoundaryConditions/DirichletAnalytic.cpp:58-117): the face is given a ghost
 * slot (neighbors entry <= -2) that the exchange never overwrites, and the
 * caller supplies the exterior state once (static solutions) or per step:
 * data host [n_slots][halo_comps][N^2] = exterior evolved variables | the
 * interior element's inverse-Jacobian row of that face (3) | interior gamma1,
 * gamma2 (GH) or gamma2 (ScalarWave).  The kernel then normalises minus the
 * interior normal with the exterior metric exactly like :545-560. */
int dgrhs_set_boundary_ghost_data(dgrhs_ctx* ctx, int slot_begin, int n_slots,
                                  const double* data);
void* dgrhs_halo_send_ptr(dgrhs_ctx* ctx);
void* dgrhs_halo_recv_ptr(dgrhs_ctx* ctx);
int dgrhs_halo_comps(dgrhs_ctx* ctx);

/* The exchange itself inside the library (one process per GPU, NCCL over NVLink):
 * replaces send_data_for_fluxes (ComputeTimeDerivative.hpp:652-774: every element
 * sends its mortar data to the neighbour's inbox) and
 * receive_boundary_data_global_time_stepping (ApplyBoundaryCorrections.hpp:205-380)
 * at batch granularity: ONE ncclSend/ncclRecv pair per peer rank and RHS carries
 * every cut mortar face, on the context's communication stream, overlapped with the
 * kernels of the elements that need no halo.
 *   dgrhs_comm_unique_id  rank 0 creates the 128-byte ncclUniqueId and hands it to
 *                         the other ranks by any means (file, MPI, torch.distributed)
 *   dgrhs_comm_init       every rank, once (collective: ncclCommInitRank)
 *   dgrhs_set_halo_peers  faces sent to / received from each rank [world]; the send
 *                         buffer (order of dgrhs_set_halo_map) and the ghost slots are
 *                         segmented by peer in rank order
 *   dgrhs_exchange_halo   pack + exchange, the context's stream waits for the halo
 *                         (callers that drive the substeps themselves)
 * After dgrhs_comm_init, dgrhs_take_steps runs the whole multi-GPU schedule itself:
 * pack -> NCCL exchange (comm stream) | interior faces + interior volume (main
 * stream) | remaining faces + boundary volume (comm stream, after the halo, next
 * to the interior volume kernel) -> join. */
int dgrhs_comm_unique_id(void* unique_id_128_bytes);
int dgrhs_comm_init(dgrhs_ctx* ctx, const void* unique_id_128_bytes, int rank, int world);
int dgrhs_set_halo_peers(dgrhs_ctx* ctx, const int32_t* send_counts,
                         const int32_t* recv_counts);
int dgrhs_exchange_halo(dgrhs_ctx* ctx);
/* Event timeline of the multi-GPU schedule (profiles/: the substitute for a per-rank
 * nsys timeline, nsys is not installed here).  When enabled, every RHS evaluation of
 * dgrhs_take_steps records CUDA events; dgrhs_get_phase_times returns, for the LAST
 * evaluation, the milliseconds from its start to the end of: pack (main stream),
 * interior faces (main), interior volume (main), the wait for the packed faces (comm
 * stream: NCCL start), NCCL send/recv (comm), remaining faces (comm), boundary volume
 * (comm) -- ms[0..6]. */
int dgrhs_set_phase_timing(dgrhs_ctx* ctx, int enable);
int dgrhs_get_phase_times(dgrhs_ctx* ctx, double* ms);

/* ---- time stepping (Time/TimeSteppers, Time/Actions) ------------------- */

/* Selects the stepper and resets time to t0, step to dt.  For Adams-Bashforth
 * of order k > 1 the history is built by the reference's forward self-start
 * (Time/Actions/SelfStartActions.hpp:181-243,316-394) on the first step. */
int dgrhs_set_stepper(dgrhs_ctx* ctx, int stepper, int order, double t0,
                      double dt);
/* Take n_steps full steps: per substep ComputeTimeDerivative,
 * ApplyBoundaryCorrections, RecordTimeStepperData, UpdateU, CleanHistory,
 * AdvanceTime (step_actions, EvolveScalarWave.hpp:229-253). */
int dgrhs_take_steps(dgrhs_ctx* ctx, int n_steps);
double dgrhs_time(dgrhs_ctx* ctx);
int64_t dgrhs_rhs_evaluations(dgrhs_ctx* ctx);
/* multi-GPU stepping in pieces: same as take_steps(1) but the caller drives
 * the RHS (so it can interleave the halo exchange).  begin_substep returns in
 * *time the time at which the RHS must be evaluated; end_substep records the
 * derivative and updates u. is_step_done is set when a full step completed. */
int dgrhs_begin_substep(dgrhs_ctx* ctx, double* time);
/* Exact time bookkeeping of the reference (Time/Slab.hpp, Time/Time.hpp:114-117,
 * TimeStepId.cpp:66-82): after dgrhs_set_stepper, declare the first slab and the number
 * of (equal) steps per slab.  Step and substep times are then formed as the reference
 * forms them -- slab k+1 = [end_k, end_k + (end_k - start_k)], time = (1 - f) start +
 * f end with the exact rational slab fraction f, substep time = (1 - c) t_step +
 * c t_next_step, step size = (end - start) * (1 / steps_per_slab) -- instead of
 * t0 + k dt.  host/SpectreTime.hpp holds the matching Slab/Time/TimeDelta/TimeStepId
 * value types; DgTimeLoop checks every substep time against the TimeStepId's. */
int dgrhs_set_slab(dgrhs_ctx* ctx, double slab_start, double slab_end, int steps_per_slab);
/* number of self-start RHS evaluations (SelfStartActions.hpp) still to come before the
 * first regular step; 0 for substep methods */
int dgrhs_self_start_substeps_left(dgrhs_ctx* ctx, int* n);
/* Non-conforming (h-refined, 2:1) mortars, equal N on both sides.
 * Faces on either side of such an interface carry DGRHS_NEIGHBOR_HANGING in the
 * neighbor table of dgrhs_set_geometry (Element<3>::neighbors() holds several
 * ids for that direction); the mortar table lists, per mortar (= fine face), the
 * row {coarse element, its direction, fine element, its direction, size_a,
 * size_b}: Spectral::MortarSize of the fine face inside the coarse face per face
 * dimension (first remaining dimension first; 0 Full, 1 LowerHalf, 2 UpperHalf),
 * i.e. dg::mortar_size(coarse, fine, dimension, orientation)
 * (NumericalAlgorithms/DiscontinuousGalerkin/MortarHelpers.cpp:51-77), given in
 * the COARSE element's logical frame.  Blocks that are not aligned
 * (OrientationMap::is_aligned() false): the fine direction is whichever face of
 * the fine element touches the coarse face, and the entry is written
 * direction | (perm << 3) with the face permutation bits of
 * dgrhs_set_neighbor_orientations, taking a mortar point (a, b) in the coarse
 * element's face frame to the fine element's face point it coincides with
 * (the orient_variables_on_slice applied to exchanged mortar data,
 * ComputeTimeDerivative.hpp:712-721).  Aligned blocks: perm = 0.
 * Semantics: InternalMortarDataImpl.hpp:230-320 (package on the face, then
 * dg::project_to_mortar, MortarHelpers.hpp:74-98) and ApplyBoundaryCorrections.hpp:
 * 797-1045 (dg_boundary_terms on the mortar, dg::project_from_mortar,
 * MortarHelpers.hpp:100-129, lift_flux with the face normal magnitude,
 * add_slice_to_data; the mortars of one coarse face are summed in table order).
 * A side that lives on another rank is written element = -(slot + 2): its face
 * (packed by the owner with dgrhs_pack_halo like every cut face) arrives in ghost
 * slot `slot`; the rank of the coarse side projects and sums the coarse
 * correction, the rank of the fine side lifts the fine one. */
#define DGRHS_NEIGHBOR_HANGING (-2147483647 - 1)
/* Neighbor-table entry of an EXTERNAL face with gh::BoundaryConditions::
 * ConstraintPreservingBjorhus, Type ConstraintPreserving (DGRHS_NEIGHBOR_BJORHUS)
 * or ConstraintPreservingPhysical (DGRHS_NEIGHBOR_BJORHUS_PHYSICAL) (GeneralizedHarmonic/
 * BoundaryConditions/Bjorhus.cpp:104-391, BjorhusImpl.cpp; a TimeDerivative-type
 * condition applied by BoundaryConditionsImpl.hpp:566-670): the corrections
 * computed from the volume time derivative, the volume partial derivatives, the
 * gauge source and the constraint fields on the face are added to dt(g, Pi,
 * Phi) on the face points.  Needs inertial coordinates (dgrhs_set_geometry) and
 * any of the gauges; static mesh. */
#define DGRHS_NEIGHBOR_BJORHUS (-2147483647)
#define DGRHS_NEIGHBOR_BJORHUS_PHYSICAL (-2147483646)
int dgrhs_set_mortars(dgrhs_ctx* ctx, int n_mortars, const int32_t* mortars);
/* p-refinement (elements with different numbers of grid points):
 * one context per N; a face whose neighbour lives in a context with another N is marked
 * DGRHS_NEIGHBOR_P_MORTAR in the neighbour table and listed here: table [n_faces][4] =
 * {element, direction, NB = the neighbour's points per dimension, the neighbour's direction
 * | permutation << 3 (as dgrhs_set_neighbor_orientations)}.  The mortar mesh has the larger
 * extents (dg::mortar_mesh, MortarHelpers.cpp:22-49); both sides package on their own face
 * mesh and project to the mortar, the correction is projected back and lifted
 * (MortarHelpers.hpp:74-129, Projection.cpp:57-362, ApplyBoundaryCorrections.hpp:286-380).
 * Before every right-hand side the neighbour's face must be in place:
 *   dgrhs_set_halo_map + dgrhs_pack_halo on the neighbour's context (its faces in halo slots),
 *   dgrhs_p_mortar_transfer(neighbour ctx, this ctx, n, halo slots, face indices of this table)
 * (stream-ordered device copies, same device), then dgrhs_compute_time_derivative_range. */
#define DGRHS_NEIGHBOR_P_MORTAR (-2147483645)
int dgrhs_set_p_mortars(dgrhs_ctx* ctx, int n_faces, const int32_t* table);
int dgrhs_p_mortar_transfer(dgrhs_ctx* src, dgrhs_ctx* dst, int n, const int32_t* src_slots,
                            const int32_t* dst_faces);
/* Spectral::projection_matrix_parent_to_child (child_to_parent = 0; Projection.cpp:
 * 279-362) / projection_matrix_child_to_parent (= 1; :57-262, operand not massive)
 * for Legendre-Gauss-Lobatto meshes with n_points_1d points on both sides,
 * row-major [target point][source point]. */
int dgrhs_projection_matrix(int n_points_1d, int child_to_parent, int size, double* matrix);
/* The same for meshes with different numbers of points (p-refinement:
 * Spectral::projection_matrix_parent_to_child(parent_mesh, child_mesh, size) /
 * projection_matrix_child_to_parent(parent_mesh, child_mesh, size), Projection.cpp:
 * 57-362; the child (mortar) mesh is the finer one, 2 <= n_parent <= n_child <= 12):
 * parent -> child [n_child][n_parent], child -> parent [n_parent][n_child],
 * row-major [target point][source point].  Host-only; the batched path itself
 * runs one N per context (no p-mortars yet). */
int dgrhs_projection_matrix_meshes(int n_parent, int n_child, int child_to_parent, int size,
                                   double* matrix);
/* gh::BoundaryConditions::DemandOutgoingCharSpeeds on every external face
 * without a ghost state (neighbor -1), GeneralizedHarmonic/BoundaryConditions/
 * DemandOutgoingCharSpeeds.cpp:37-76 (applied by BoundaryConditionsImpl.hpp:
 * 672-764 for Type::DemandOutgoingCharSpeeds): no boundary correction is added
 * on such a face, and every RHS evaluation checks that the four characteristic
 * speeds (gh::characteristic_speeds, Characteristics.cpp:24-40) with respect to
 * the outward unit normal are non-negative at every face point.  The reference
 * ERRORs at once; here the violation is latched on the device and
 * dgrhs_check_outgoing_char_speeds returns non-zero (with the reference's
 * message in dgrhs_last_error) at the caller's next check.  enable = 0 is the
 * reference's `None`-like behaviour (no correction, no check). */
int dgrhs_set_demand_outgoing_char_speeds(dgrhs_ctx* ctx, int enable);
int dgrhs_check_outgoing_char_speeds(dgrhs_ctx* ctx, long long* n_violations,
                                     double* min_speed);
/* UpdateU (Time/Actions/UpdateU.hpp:82-89) is fused into the volume kernel by
 * default: u_new = a*u + sum_j c_j v_j, same coefficients and term order as the
 * separate update, written to a second state buffer that becomes the state at
 * end_substep (so dgrhs_state_device_ptr changes from step to step).
 * enable = 0 selects the separate update kernel (bit-identical results). */
int dgrhs_set_fused_update(dgrhs_ctx* ctx, int enable);
/* dg::Actions::Filter<Filters::Exponential<0>> after UpdateU (ParallelAlgorithms/
 * Actions/FilterAction.hpp, NumericalAlgorithms/LinearOperators/
 * ExponentialFilter.cpp:45-76; KerrSchild.yaml:127-132 uses Alpha 36, HalfPower
 * 64): every evolved component of every element is multiplied by the filter
 * matrix along xi, eta, zeta after each substep update.  Disabled by default. */
int dgrhs_set_exponential_filter(dgrhs_ctx* ctx, int enable, double alpha, int half_power);
/* Filters::Exponential applied once to the resident state, outside a substep:
 * apply_matrices(u, {F, F, F}) on every evolved component (LinearOperators/
 * ExponentialFilter.cpp:45-76; the action dg::Actions::Filter, Filtering.hpp:120-165,
 * runs it after every substep update -- dgrhs_end_substep does that itself). */
int dgrhs_apply_exponential_filter(dgrhs_ctx* ctx);
/* Spectral::filtering::exponential_filter(Mesh<1>{N, Legendre, GaussLobatto},
 * alpha, half_power) (Spectral/Filtering.cpp:20-32), row-major [N*N]. */
int dgrhs_exponential_filter_matrix(int n_points_1d, double alpha, int half_power,
                                    double* matrix);
/* GH volume work as two kernels (pointwise context kernel + high-occupancy
 * streaming kernel at 16 warps/SM; opt-in, N <= 10) instead of the single
 * fused kernel (default).  Results agree to rounding.  Measured on B200
 * (config 2): 0.60 + 1.18 ms vs 1.74 ms for the fused kernel -- the streaming
 * kernel becomes shared-memory bound (70 % of LSU wavefront peak), see
 * profiles/. */
int dgrhs_set_split_volume(dgrhs_ctx* ctx, int enable);
int dgrhs_end_substep(dgrhs_ctx* ctx, int* is_step_done);

/* Measurement aid for bench.py (roofline of the individual kernels): runs the
 * face kernel, the volume kernel and a k-term stepper update `reps` times
 * each on the context's stream, bracketed by CUDA events, and returns the mean
 * milliseconds per launch in ms[0..2]; ms[3] is the volume kernel with the
 * stepper update fused in (the kernel the steppers actually use); ms[4] is the
 * exponential filter pass over a scratch copy of the state (0 if the filter is
 * not enabled).  ms must hold 5 doubles.  Does not change u (updates and the
 * filter go to scratch buffers). */
int dgrhs_time_kernels(dgrhs_ctx* ctx, int reps, int update_terms, double* ms);

/* GH constraint diagnostics of the current state, as the
 * L2Norm with Components: Sum of ObserveNorms (ParallelAlgorithms/Events/
 * ObserveNorms.hpp:60-80) over this context's elements: norms[0] gauge
 * constraint C_a = H_a + Gamma_a (GeneralizedHarmonic/Constraints.cpp:965-1000),
 * norms[1] three-index constraint d_i g_ab - Phi_iab (:935-962), norms[2]
 * four-index constraint eps_ijk d_j Phi_kab (:1070-1100). */
int dgrhs_gh_constraint_norms(dgrhs_ctx* ctx, double* norms);

/* Synchronise the context's stream. */
int dgrhs_synchronize(dgrhs_ctx* ctx);
/* cudaStream_t used by the context (for CUDA-event timing by the caller). */
void* dgrhs_stream(dgrhs_ctx* ctx);
/* Device pointer to the evolved variables [n_elements][n_vars][n_padded] and
 * the padded per-component stride. */
void* dgrhs_state_device_ptr(dgrhs_ctx* ctx);
int dgrhs_padded_points(dgrhs_ctx* ctx);

/* ---- single-operator entry points (host pointers, one element) --------- *
 * Thin GPU forwards for the reference's operator surface, used by the C++
 * shims in spectre_b200/host/ and by the parity tests.                      */

/* partial_derivatives (LinearOperators/PartialDerivatives.tpp:191-241):
 * u [n_comps][n], inv_jacobian [9][n] -> du [3*n_comps][n], d_i u_c at 3c+i */
int dgrhs_partial_derivatives(int n_points_1d, int n_comps, const double* u,
                              const double* inv_jacobian, double* du);
/* gh::TimeDerivative<3>::apply (GeneralizedHarmonic/TimeDerivative.hpp:143-192,
 * TimeDerivative.cpp:31-407) for n points of one element: u [50][n] (g, Pi,
 * Phi), du [150][n] (d_i of component c at 3c+i, i.e. d_spacetime_metric,
 * d_pi, d_phi), gamma0/1/2 [n]; harmonic != 0 selects gauges::Harmonic,
 * otherwise H_a [4][n] and d_a H_b [16][n] (a + 4b) are the output of
 * gauges::dispatch.  Output dt_u [50][n].  The 29 temporary tensors of the
 * reference signature are not materialised. */
int dgrhs_gh_time_derivative(int n, const double* u, const double* du,
                             const double* gamma0, const double* gamma1,
                             const double* gamma2, int harmonic,
                             const double* gauge_h, const double* d4_gauge_h,
                             double* dt_u);
/* gh::BoundaryConditions::ConstraintPreservingBjorhus<3>::dg_time_derivative
 * (GeneralizedHarmonic/BoundaryConditions/Bjorhus.hpp, Bjorhus.cpp:104-391) on n
 * face points of a static mesh (face_mesh_velocity == nullopt), arguments in the
 * reference's order and Tensor storage order ([independent component][n]:
 * tnsr::i 3, tnsr::aa / AA 10, tnsr::iaa 30 at i + 3 sym, tnsr::ijaa 90 at
 * i + 3 (j + 3 sym), tnsr::ab 16 at a + 4 b); physical = 0 Type
 * ConstraintPreserving, 1 ConstraintPreservingPhysical.  The unused
 * normal_vector and d_spacetime_metric arguments of the reference are omitted. */
int dgrhs_gh_bjorhus_dg_time_derivative(
    int n, int physical, const double* normal_covector, const double* spacetime_metric,
    const double* pi, const double* phi, const double* coords, const double* gamma1,
    const double* gamma2, const double* lapse, const double* shift,
    const double* inverse_spacetime_metric, const double* spacetime_unit_normal_vector,
    const double* three_index_constraint, const double* gauge_source,
    const double* spacetime_deriv_gauge_source, const double* dt_spacetime_metric,
    const double* dt_pi, const double* dt_phi, const double* d_pi, const double* d_phi,
    double* dt_spacetime_metric_correction, double* dt_pi_correction,
    double* dt_phi_correction);
/* ScalarWave::TimeDerivative<3>::apply (ScalarWave/TimeDerivative.hpp:26-50,
 * TimeDerivative.cpp:14-45): u [5][n], du [15][n], gamma2 [n] -> dt_u [5][n] */
int dgrhs_sw_time_derivative(int n, const double* u, const double* du,
                             const double* gamma2, double* dt_u);
/* gh::BoundaryCorrections::UpwindPenalty<3>::dg_package_data
 * (UpwindPenalty.hpp:226-262, UpwindPenalty.cpp:36-158) on f face points:
 * packaged [134][f] in dg_package_field_tags order (v_g 10 | v_zero 30 |
 * v_plus 10 | v_minus 10 | n v_plus 30 | n v_minus 30 | gamma2 v_g 10 |
 * speeds 4); returns the max char speed like the reference. */
int dgrhs_gh_package_data(int f, const double* u, const double* gamma1,
                          const double* gamma2, const double* lapse,
                          const double* shift, const double* normal_covector,
                          const double* normal_vector, double* packaged,
                          double* max_abs_char_speed);
/* The same with normal_dot_mesh_velocity [f] of a moving mesh (UpwindPenalty.cpp:85-91: the
 * speeds relative to the mesh; NULL = static mesh). */
int dgrhs_gh_package_data_moving(int f, const double* u, const double* gamma1,
                                 const double* gamma2, const double* lapse,
                                 const double* shift, const double* normal_covector,
                                 const double* normal_vector,
                                 const double* normal_dot_mesh_velocity, double* packaged,
                                 double* max_abs_char_speed);
/* ...::dg_boundary_terms (UpwindPenalty.cpp:161-275): [134][f] x2 -> [50][f] */
int dgrhs_gh_boundary_terms(int f, const double* packaged_int,
                            const double* packaged_ext,
                            double* boundary_correction);
/* ScalarWave::BoundaryCorrections::UpwindPenalty<3> (UpwindPenalty.hpp:211-260,
 * UpwindPenalty.cpp:36-205): packaged [16][f], correction [5][f] */
int dgrhs_sw_package_data(int f, const double* u, const double* gamma2,
                          const double* normal_covector, double* packaged,
                          double* max_abs_char_speed);
/* with normal_dot_mesh_velocity [f] (ScalarWave UpwindPenalty.cpp:55-67; NULL = static) */
int dgrhs_sw_package_data_moving(int f, const double* u, const double* gamma2,
                                 const double* normal_covector,
                                 const double* normal_dot_mesh_velocity, double* packaged,
                                 double* max_abs_char_speed);
int dgrhs_sw_boundary_terms(int f, const double* packaged_int,
                            const double* packaged_ext,
                            double* boundary_correction);
/* dg::lift_flux (NumericalAlgorithms/DiscontinuousGalerkin/LiftFlux.hpp:41-62),
 * in place on [n_comps][f] */
int dgrhs_lift_flux(int f, int n_comps, double* boundary_correction,
                    int extent_perpendicular_to_boundary,
                    const double* magnitude_of_face_normal);

/* Spectral::differentiation_matrix(Mesh<1>{N, Legendre, GaussLobatto})
 * (Spectral.cpp:431-445), row-major D[i*N + j]; collocation points/weights
 * (Legendre.cpp:187-232). */
int dgrhs_differentiation_matrix(int n_points_1d, double* matrix);
int dgrhs_collocation_points_and_weights(int n_points_1d, double* points,
                                         double* weights);
/* adams_coefficients::coefficients (AdamsCoefficients.hpp:64-104,
 * AdamsCoefficients.cpp:13-42,75-117): history times oldest first. */
int dgrhs_adams_bashforth_coefficients(int order, const double* history_times,
                                       double step_start, double step_end,
                                       double* coefficients);

/* TimeStepper::order(), number_of_substeps(), number_of_past_steps(),
 * stable_step() (Time/TimeSteppers/TimeStepper.hpp:47-246; AdamsBashforth.cpp:60-95,
 * Rk3HesthavenSsp.cpp:21-26, Rk3Owren.cpp:8-15, Rk3Kennedy.cpp:8-10,
 * ClassicalRungeKutta4.cpp:10-22, DormandPrince5.cpp:8-19).  `order` is read for
 * AdamsBashforth only; any output pointer may be NULL.  Host-only. */
int dgrhs_stepper_properties(int stepper, int order, int* order_out, int* number_of_substeps,
                             int* number_of_past_steps, double* stable_step);

/* Row `substep` of the Butcher tableau as RungeKutta::update_u_impl uses it
 * (RungeKutta.cpp:69-122): the coefficients of f_0 .. f_substep for the update that
 * follows substep `substep`; the last substep returns the result coefficients. */
int dgrhs_butcher_row(int stepper, int substep, double* coefficients);

/* TimeStepper::update_u(u, history, time_step) (TimeStepper.hpp:96-102) on a flat span of
 * `size` doubles (host pointers, one round trip).  history_derivatives [n_history][size],
 * oldest first.  AdamsBashforth (AdamsBashforth.cpp:120-135): n_history = order entries at
 * history_times (arbitrary spacing, AdamsCoefficients.hpp:64-104), u holds the value at the
 * newest time and becomes the value one time_step later.  Substep methods: the history holds
 * the derivatives of the substeps done so far in this step, step_start_value the value at
 * the start of the step; u holds the current substep value and becomes the next one
 * (Rk3HesthavenSsp.cpp:63-81, RungeKutta.cpp:69-122). */
int dgrhs_update_u(int stepper, int order, long long size, double* u, int n_history,
                   const double* history_times, const double* history_derivatives,
                   const double* step_start_value, double time_step);

/* dg::project_to_mortar / dg::project_from_mortar (NumericalAlgorithms/DiscontinuousGalerkin/
 * MortarHelpers.hpp:74-129) for variables on a face: [n_comps][extent_b][extent_a], a fastest.
 * face_extents / mortar_extents: points per face dimension (mortar >= face,
 * MortarHelpers.cpp:22-49); mortar_size: DGRHS_MORTAR_* per dimension.  To the mortar:
 * interpolation (Projection.cpp:279-362); from the mortar: L2 projection
 * (Projection.cpp:57-262).  Dimensions that need no projection are skipped like
 * apply_matrices does. */
int dgrhs_project_to_mortar(int n_comps, const int* face_extents, const int* mortar_extents,
                            const int* mortar_size, const double* face_vars, double* mortar_vars);
int dgrhs_project_from_mortar(int n_comps, const int* face_extents, const int* mortar_extents,
                              const int* mortar_size, const double* mortar_vars,
                              double* face_vars);

/* orient_variables_on_slice (Domain/Structure/OrientationMapHelpers.cpp:25-120) for the face
 * of a 3-D element: vars [n_comps][extent_b][extent_a] in this element's frame ->
 * the neighbour's frame.  permutation as in dgrhs_set_neighbor_orientations: bit 0 swaps the
 * two face coordinates, bits 1 / 2 flip the neighbour's first / second face coordinate
 * (host/SpectreShims.hpp face_orientation() derives it from an OrientationMap<3>). */
int dgrhs_orient_variables_on_slice(int n_comps, const int* slice_extents, int permutation,
                                    const double* vars, double* oriented);

/* Fraction of the step at which substep k = 1 .. number_of_substeps-1 is evaluated
 * (TimeStepper::next_time_id: RungeKutta.cpp:34-58 butcher_tableau().substep_times,
 * Rk3HesthavenSsp.cpp:36-48 {1, 1/2}; AdamsBashforth has no substeps).  Host-only. */
int dgrhs_stepper_substep_fractions(int stepper, double* fractions);

#ifdef __cplusplus
}
#endif
#endif /* DGRHS_H */
