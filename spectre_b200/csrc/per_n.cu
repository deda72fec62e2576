// Every kernel instantiation of ONE number of grid points per dimension
// (compile with -DDG_N=<2..12>; __graft_entry__.build() compiles the eleven
// copies in parallel) and the launchers that queue them, exported to dgrhs.cu as
// a table (DgNOps, ctx.cuh).
#ifndef DG_N
#error "compile with -DDG_N=<points per dimension>"
#endif

#include "ctx.cuh"

namespace {

// ConstraintPreservingBjorhus faces (of the elements [eb, ee) if ee > 0)
template <int N>
int launch_bjorhus(dgrhs_ctx* c, cudaStream_t stream, int eb, int ee) {
  dg::BjorhusArgs b{c->u, c->invjac, c->stat, c->gH, c->gdH, c->coords, c->D, c->corr,
                    c->bjorhus_faces, {}, eb, ee};
  int gauge_mode = 1;
  if (c->gauge == DGRHS_GAUGE_HARMONIC) gauge_mode = 0;
  if (c->gauge == DGRHS_GAUGE_DAMPED_HARMONIC) {
    const double* p = c->gauge_params;
    b.dh = {p[0], p[1], p[2], p[3], (int)p[4], (int)p[5], (int)p[6]};
    gauge_mode = 2;
  }
  constexpr int bT = (N * N + 31) / 32 * 32;
  dg::gh_bjorhus_kernel<N><<<c->n_bjorhus_faces, bT, 0, stream>>>(b, gauge_mode);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  return 0;
}

template <int N>
int launch_faces(dgrhs_ctx* c, int eb, int ee) {
  if (ee <= eb) return 0;
  if (!c->nbr_face && !c->aligned_table_ok)
    return fail("neighbor table is not that of aligned blocks: call "
                "dgrhs_set_neighbor_orientations");
  // whole batch: every interface once; element ranges: the interior / boundary
  // split of the multi-GPU schedule (see FaceArgs::pass)
  int pass = 0, n_int = c->nelem;
  if (!(eb == 0 && ee == c->nelem)) {
    if (c->n_interior < 0)
      return fail("element ranges need dgrhs_set_interior_count (interior elements first)");
    n_int = c->n_interior;
    if (eb == 0 && ee == n_int)
      pass = 1;
    else if (eb == n_int && ee == c->nelem)
      pass = 2;
    else
      return fail("element range must be [0, n_interior) or [n_interior, n_elements)");
  }
  dg::FaceArgs a{c->u,    c->invjac, c->stat, c->nbr, c->nbr_face, c->halo_recv,
                 c->corr, c->nelem,  n_int,   pass,   pass == 2 ? n_int : 0, c->violations,
                 c->mesh_v};
  if (c->mesh_v && (c->n_bjorhus_faces > 0 || c->n_mortar_faces > 0 || c->n_pmortar_faces > 0))
    return fail("moving mesh: Bjorhus faces and non-conforming mortars are not supported");
  // Bjorhus faces and non-conforming mortars need no halo data: all of them are
  // evaluated once per right-hand side, with whichever pass comes first, so that
  // their corrections are in place before ANY volume kernel of this evaluation
  const bool aux_now = c->aux_faces_eval != c->rhs_evals;
  c->aux_faces_eval = c->rhs_evals;
  const bool bjorhus_now = c->n_bjorhus_faces > 0 && aux_now;
  if (bjorhus_now) {
    CU(cudaEventRecord(c->aux_fork, c->stream));
    CU(cudaStreamWaitEvent(c->aux_stream, c->aux_fork, 0));
    if (launch_bjorhus<N>(c, c->aux_stream, 0, 0)) return 1;
    CU(cudaEventRecord(c->aux_join, c->aux_stream));
  }
  const long long total = (long long)(c->nelem - a.elem_begin) * 6 * N * N;
  const int blocks = (int)((total + 127) / 128);
  if (c->system == DGRHS_SYSTEM_GH)
    dg::gh_face_kernel<N><<<blocks, 128, 0, c->stream>>>(a);
  else
    dg::sw_face_kernel<N><<<blocks, 128, 0, c->stream>>>(a);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  // whole-batch pass of the fused GH kernel: the join is deferred to the volume launch of the
  // elements that own a Bjorhus face (rhs_range); otherwise join here
  const bool defer_join = bjorhus_now && pass == 0 && c->bjorhus_tail_begin > 0 &&
                          c->volume_variant == 0 &&
                          std::getenv("DGRHS_NO_BJORHUS_OVERLAP") == nullptr;
  if (bjorhus_now && !defer_join) CU(cudaStreamWaitEvent(c->stream, c->aux_join, 0));
  c->bjorhus_join_pending = defer_join;
  // mortar groups whose sides are all local run with the first pass; groups with a
  // remote side need the halo: with the boundary pass (or the single full pass)
  auto launch_mortars = [&](int first, int count) -> int {
    if (count <= 0) return 0;
    dg::MortarArgs m{c->u, c->invjac, c->stat, c->corr, c->mortar_faces, c->mortar_table,
                     c->mortar_P, c->mortar_R, c->halo_recv, first};
    constexpr int msmem = dg::mortar_smem_bytes<N>();
    constexpr int mT = (N * N + 31) / 32 * 32;
    if (c->system == DGRHS_SYSTEM_GH) {
      auto k = dg::mortar_kernel<N, 1>;
      CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, msmem));
      k<<<count, mT, msmem, c->stream>>>(m);
    } else {
      auto k = dg::mortar_kernel<N, 0>;
      CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, msmem));
      k<<<count, mT, msmem, c->stream>>>(m);
    }
    dgrhs_internal_count_launch();
    CU(cudaGetLastError());
    return 0;
  };
  if (aux_now && launch_mortars(0, c->n_mortar_faces_local)) return 1;
  if (pass != 1 &&
      launch_mortars(c->n_mortar_faces_local, c->n_mortar_faces - c->n_mortar_faces_local))
    return 1;
  // faces to neighbours with another N: the neighbour's face was transferred before this
  // right-hand side (dgrhs_p_mortar_transfer); with the pass that has the halo
  if (pass != 1 && c->n_pmortar_faces > 0) {
    dg::PMortarArgs m{c->u, c->invjac, c->stat, c->corr, c->pm_faces, c->pm_ghost, c->pm_P, c->pm_R};
    constexpr int psmem = dg::pmortar_smem_bytes();
    if (c->system == DGRHS_SYSTEM_GH) {
      auto k = dg::pmortar_kernel<N, 1>;
      CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, psmem));
      k<<<c->n_pmortar_faces, dg::kPMortarThreads, psmem, c->stream>>>(m);
    } else {
      auto k = dg::pmortar_kernel<N, 0>;
      CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, psmem));
      k<<<c->n_pmortar_faces, dg::kPMortarThreads, psmem, c->stream>>>(m);
    }
    dgrhs_internal_count_launch();
    CU(cudaGetLastError());
  }
  c->pdl_volume = g_pdl && (!bjorhus_now || defer_join) && c->n_mortar_faces == 0 &&
                  c->n_pmortar_faces == 0;
  return 0;
}

template <int N>
int launch_gauge(dgrhs_ctx* c, double time) {
  if (c->gauge != DGRHS_GAUGE_ANALYTIC_GAUGE_WAVE) return 0;
  if (!c->coords) return fail("AnalyticChristoffel(GaugeWave) gauge needs coordinates");
  dg::GaugeWaveArgs g{c->coords, c->gH, c->gauge_params[0], c->gauge_params[1], time,
                      c->nelem};
  const long long total = (long long)c->nelem * c->n;
  dg::gauge_wave_h_kernel<N><<<(int)((total + 255) / 256), 256, 0, c->stream>>>(g);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  // spatial derivative d_i H_b -> gdH component (i+1) + 4 b; d_0 H_b stays 0
  dg::DerivArgs d{c->gH, c->invjac, c->gdH, c->D, 4, 16, 1, 4};
  dg::partial_derivatives_kernel<N><<<c->nelem * 4, 256, 0, c->stream>>>(d);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  return 0;
}

template <int N, int kGauge>
int launch_gh_split(dgrhs_ctx* c, const dg::GhVolArgs& a, int eb, int ee) {
  if (!c->ctxbuf &&
      dev_alloc(&c->ctxbuf, (size_t)c->nelem * dg::kGhCtxComps * c->npad))
    return 1;
  dg::GhCtxArgs ca{c->u, c->stat, c->gH, c->gdH, c->coords, c->ctxbuf, a.dh, eb, ee};
  const long long pts = (long long)(ee - eb) * c->n;
  dg::gh_context_kernel<N, kGauge><<<(int)((pts + 255) / 256), 256, 0, c->stream>>>(ca);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  constexpr int smem = dg::SCfg<N>::smem_bytes;
  auto k = dg::gh_stream_kernel<N>;
  CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k<<<(ee - eb) * dg::Cfg<N>::nchunk, dg::Cfg<N>::T, smem, c->stream>>>(a, c->ctxbuf);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  return 0;
}

template <int N>
int launch_volume(dgrhs_ctx* c, double* dt, int eb, int ee, bool with_corr,
                  const dg::UpdateArgs& upd = dg::UpdateArgs{}) {
  const bool pdl = c->pdl_volume;
  c->pdl_volume = false;
  if (ee <= eb) return 0;
  const int blocks = (ee - eb) * dg::Cfg<N>::nchunk;
  if (c->system == DGRHS_SYSTEM_GH) {
    dg::GhVolArgs a{c->u, dt, c->invjac, c->stat, with_corr ? c->corr : nullptr,
                    c->gH, c->gdH, c->D, c->coords, {}, eb, upd, 0};
    {
      // one CTA per SM (N >= 10): pull the inputs of the CTA one wave ahead into L2
      static const char* env = std::getenv("DGRHS_PREFETCH_DIST");
      a.prefetch_dist = env ? std::atoi(env)
                            : (dg::Cfg<N>::two_cta ? 0 : c->num_sms * dg::Cfg<N>::min_blocks);
    }
    if constexpr (dg::SCfg<N>::fits && N <= 10) {
      if (c->volume_variant == 1) {
        if (c->gauge == DGRHS_GAUGE_HARMONIC) return launch_gh_split<N, 0>(c, a, eb, ee);
        if (c->gauge == DGRHS_GAUGE_DAMPED_HARMONIC) {
          if (!c->coords) return fail("DampedHarmonic gauge needs inertial coordinates");
          const double* p = c->gauge_params;
          a.dh = {p[0], p[1], p[2], p[3], (int)p[4], (int)p[5], (int)p[6]};
          return launch_gh_split<N, 2>(c, a, eb, ee);
        }
        return launch_gh_split<N, 1>(c, a, eb, ee);
      }
    }
    if (c->gauge == DGRHS_GAUGE_DAMPED_HARMONIC) {
      if (!c->coords) return fail("DampedHarmonic gauge needs inertial coordinates");
      const double* p = c->gauge_params;
      a.dh = {p[0], p[1], p[2], p[3], (int)p[4], (int)p[5], (int)p[6]};
    }
    constexpr int smem = dg::gh_volume_smem_bytes<N>();
    if (c->gauge == DGRHS_GAUGE_HARMONIC) {
      auto k = dg::gh_volume_kernel<N, 0>;
      CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CU(launch_dependent(k, blocks, dg::Cfg<N>::T, smem, c->stream, pdl, a));
    } else if (c->gauge == DGRHS_GAUGE_DAMPED_HARMONIC) {
      if (!c->coords) return fail("DampedHarmonic gauge needs inertial coordinates");
      const double* p = c->gauge_params;
      a.dh = {p[0], p[1], p[2], p[3], (int)p[4], (int)p[5], (int)p[6]};
      auto k = dg::gh_volume_kernel<N, 2>;
      CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CU(launch_dependent(k, blocks, dg::Cfg<N>::T, smem, c->stream, pdl, a));
    } else {
      auto k = dg::gh_volume_kernel<N, 1>;
      CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CU(launch_dependent(k, blocks, dg::Cfg<N>::T, smem, c->stream, pdl, a));
    }
  } else {
    dg::SwVolArgs a{c->u, dt, c->invjac, c->stat, with_corr ? c->corr : nullptr, c->D, eb,
                    upd};
    constexpr int smem = dg::sw_volume_smem_bytes<N>();
    auto k = dg::sw_volume_kernel<N>;
    CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k<<<blocks, dg::Cfg<N>::T, smem, c->stream>>>(a);
  }
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  return 0;
}

template <int N>
int launch_pack(dgrhs_ctx* c) {
  if (c->n_send == 0) return 0;
  dg::PackArgs a{c->u, c->invjac, c->stat, c->halo_map, c->halo_send, c->n_send};
  const long long total = (long long)c->n_send * N * N;
  const int blocks = (int)((total + 127) / 128);
  if (c->system == DGRHS_SYSTEM_GH)
    dg::pack_halo_kernel<N, 50><<<blocks, 128, 0, c->stream>>>(a);
  else
    dg::pack_halo_kernel<N, 5><<<blocks, 128, 0, c->stream>>>(a);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  return 0;
}

template <int N>
int launch_volume_p(dgrhs_ctx* c, double* dt, int eb, int ee, bool with_corr,
                    const dg::UpdateArgs* upd) {
  return launch_volume<N>(c, dt, eb, ee, with_corr, upd ? *upd : dg::UpdateArgs{});
}

template <int N>
int launch_filter_range(dgrhs_ctx* c, double* u, int ntiles) {
  if (ntiles <= 0) return 0;
  dg::FilterArgs a;
  a.u = u;
  a.ntiles = ntiles;
  for (int k = 0; k < N * N; ++k) a.Fm[k] = c->filterF_host[k];
  using F = dg::FilterCfg<N>;
  const int ngroups = (a.ntiles + F::G - 1) / F::G;
  CU(cudaFuncSetAttribute(dg::exponential_filter_kernel<N>,
                          cudaFuncAttributeMaxDynamicSharedMemorySize, F::smem_bytes));
  const int blocks = ngroups;
  dg::exponential_filter_kernel<N><<<blocks, F::T, F::smem_bytes, c->stream>>>(a);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  return 0;
}

template <int N>
int launch_filter(dgrhs_ctx* c) {
  return launch_filter_range<N>(c, c->u, c->nelem * c->C);
}

// H_a = -Gamma_a of a given state and its spatial derivatives (AnalyticChristoffel
// gauge of a static solution, AnalyticChristoffel.cpp:76-147)
template <int N>
int launch_gauge_from_state(dgrhs_ctx* c, const double* state_dev) {
  dg::GaugeFromStateArgs g{state_dev, c->gH, c->nelem};
  const long long total = (long long)c->nelem * c->n;
  dg::gauge_h_from_state_kernel<N><<<(int)((total + 255) / 256), 256, 0, c->stream>>>(g);
  dgrhs_internal_count_launch();
  dg::DerivArgs d{c->gH, c->invjac, c->gdH, c->D, 4, 16, 1, 4};
  dg::partial_derivatives_kernel<N><<<c->nelem * 4, 256, 0, c->stream>>>(d);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  return 0;
}

template <int N>
int launch_constraints(dgrhs_ctx* c, double* sums_dev) {
  dg::ConstraintArgs a{c->u, c->invjac, c->gauge == DGRHS_GAUGE_HARMONIC ? nullptr : c->gH,
                       c->D, sums_dev};
  constexpr int smem = (4 * dg::Cfg<N>::npad + N * N) * 8;
  auto k = dg::gh_constraints_kernel<N>;
  CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k<<<c->nelem, 256, smem, c->stream>>>(a);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  return 0;
}

template <int N>
int launch_partial_derivatives(const dg::DerivArgs* a, int blocks, cudaStream_t stream) {
  dg::partial_derivatives_kernel<N><<<blocks, 256, 0, stream>>>(*a);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  return 0;
}

// moving-mesh terms on top of the static-mesh volume kernel's output
template <int N>
int launch_mesh_velocity_terms(dgrhs_ctx* c, double* dt, int eb, int ee) {
  if (ee <= eb) return 0;
  dg::MeshVelocityArgs a{c->u, c->invjac, c->stat, c->mesh_v, c->D, dt, c->C, eb,
                         c->system == DGRHS_SYSTEM_GH ? 1 : 0};
  dg::mesh_velocity_terms_kernel<N><<<(ee - eb) * c->C, 256, 0, c->stream>>>(a);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  return 0;
}

// LTS: the "volume" part of the time derivative in the reference's sense -- volume terms
// plus the external boundary conditions (ComputeTimeDerivative applies them) -- of the
// elements [eb, ee): the face kernel runs on a neighbour table whose internal faces are
// marked "no correction"
template <int N>
int launch_lts_evaluate(dgrhs_ctx* c, const int32_t* nbr_external, const uint8_t* mortar_skip,
                        double* dt, int eb, int ee, const dg::UpdateArgs* upd) {
  if (ee <= eb) return 0;
  dg::FaceArgs a{c->u, c->invjac, c->stat, nbr_external, c->nbr_face, c->halo_recv,
                 c->corr, ee, ee, 0, eb, c->violations, nullptr};
  if (c->n_bjorhus_faces > 0 && launch_bjorhus<N>(c, c->stream, eb, ee)) return 1;
  const long long total = (long long)(ee - eb) * 6 * N * N;
  const int blocks = (int)((total + 127) / 128);
  if (c->system == DGRHS_SYSTEM_GH)
    dg::gh_face_kernel<N><<<blocks, 128, 0, c->stream>>>(a);
  else
    dg::sw_face_kernel<N><<<blocks, 128, 0, c->stream>>>(a);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  // non-conforming mortars that are not in the boundary histories (both sides on this level)
  if (c->n_mortar_faces > 0) {
    dg::MortarArgs m{c->u, c->invjac, c->stat, c->corr, c->mortar_faces, c->mortar_table,
                     c->mortar_P, c->mortar_R, c->halo_recv, 0, mortar_skip, eb, ee};
    constexpr int msmem = dg::mortar_smem_bytes<N>();
    constexpr int mT = (N * N + 31) / 32 * 32;
    if (c->system == DGRHS_SYSTEM_GH) {
      auto k = dg::mortar_kernel<N, 1>;
      CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, msmem));
      k<<<c->n_mortar_faces, mT, msmem, c->stream>>>(m);
    } else {
      auto k = dg::mortar_kernel<N, 0>;
      CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, msmem));
      k<<<c->n_mortar_faces, mT, msmem, c->stream>>>(m);
    }
    dgrhs_internal_count_launch();
    CU(cudaGetLastError());
  }
  c->pdl_volume = false;
  return launch_volume<N>(c, dt, eb, ee, true, upd ? *upd : dg::UpdateArgs{});
}

template <int N>
int launch_lts_snapshot(dgrhs_ctx* c, double* fh, const uint8_t* in_history, int depth, int slot,
                        int eb, int ee) {
  if (ee <= eb) return 0;
  const long long total = (long long)(ee - eb) * 6 * N * N;
  const int blocks = (int)((total + 127) / 128);
  if (c->system == DGRHS_SYSTEM_GH)
    dg::lts_snapshot_kernel<N, 50><<<blocks, 128, 0, c->stream>>>(c->u, fh, in_history, depth,
                                                                  slot, eb, ee);
  else
    dg::lts_snapshot_kernel<N, 5><<<blocks, 128, 0, c->stream>>>(c->u, fh, in_history, depth,
                                                                 slot, eb, ee);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  return 0;
}

template <int N>
int launch_lts_boundary(dgrhs_ctx* c, const dg::LtsBoundaryArgs* a) {
  const int ne = a->elem_end - a->elem_begin;
  if (ne <= 0) return 0;
  if (a->terms) {   // conforming faces in the histories (nullptr: only the final add)
    const long long total = (long long)ne * 6 * N * N;
    const int blocks = (int)((total + 127) / 128);
    if (c->system == DGRHS_SYSTEM_GH)
      dg::gh_lts_boundary_kernel<N><<<blocks, 128, 0, c->stream>>>(*a);
    else
      dg::sw_lts_boundary_kernel<N><<<blocks, 128, 0, c->stream>>>(*a);
    dgrhs_internal_count_launch();
    CU(cudaGetLastError());
    return 0;
  }
  const long long pts = (long long)ne * c->n;
  dg::lts_add_kernel<N><<<(int)((pts + 255) / 256), 256, 0, c->stream>>>(
      a->u, a->acc, a->in_history, c->C, a->elem_begin, a->elem_end);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  return 0;
}

template <int N>
int launch_lts_mortar(dgrhs_ctx* c, const dg::LtsMortarArgs* a, int n_groups) {
  if (n_groups <= 0) return 0;
  constexpr int msmem = dg::mortar_smem_bytes<N>();
  constexpr int mT = (N * N + 31) / 32 * 32;
  if (c->system == DGRHS_SYSTEM_GH) {
    auto k = dg::lts_mortar_kernel<N, 1>;
    CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, msmem));
    k<<<n_groups, mT, msmem, c->stream>>>(*a);
  } else {
    auto k = dg::lts_mortar_kernel<N, 0>;
    CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, msmem));
    k<<<n_groups, mT, msmem, c->stream>>>(*a);
  }
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  return 0;
}

static const DgNOps kOps = {launch_faces<DG_N>,
                     launch_gauge<DG_N>,
                     launch_volume_p<DG_N>,
                     launch_pack<DG_N>,
                     launch_filter<DG_N>,
                     launch_gauge_from_state<DG_N>,
                     launch_constraints<DG_N>,
                     launch_partial_derivatives<DG_N>,
                     launch_mesh_velocity_terms<DG_N>,
                     launch_lts_evaluate<DG_N>,
                     launch_lts_snapshot<DG_N>,
                     launch_lts_boundary<DG_N>,
                     launch_lts_mortar<DG_N>,
                     launch_filter_range<DG_N>};

}  // namespace

#define DG_CAT2(a, b) a##b
#define DG_CAT(a, b) DG_CAT2(a, b)
extern "C" const DgNOps* DG_CAT(dgrhs_internal_nops_, DG_N)(void) { return &kOps; }
