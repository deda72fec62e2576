"""CPU restatement (numpy, one grid point at a time) of the constraint-preserving
Bjorhus boundary condition of the GeneralizedHarmonic system, type
`ConstraintPreserving` and `ConstraintPreservingPhysical` -- TEST INFRASTRUCTURE
ONLY (see oracle/oracle.py).

Follows the reference (sxs-collaboration/spectre v2024.09.29):
  GeneralizedHarmonic/BoundaryConditions/Bjorhus.cpp:104-391 (dg_time_derivative),
    :393-545 (compute_intermediate_vars)
  GeneralizedHarmonic/BoundaryConditions/BjorhusImpl.cpp:26-47 (dt v_psi),
    :49-103 (dt v_zero), :105-151 (gauge Sommerfeld terms), :153-221
    (constraint-dependent terms), :496-532 (dt v_minus)
  GeneralizedHarmonic/Constraints.hpp (Eq. 44 two-index constraint, Eq. 43 F
    constraint of Lindblom et al. 2005, as documented there) and Constraints.cpp
    (two_index_constraint: 11 terms from :25; f_constraint: 25 terms from :282)
  GeneralizedHarmonic/Characteristics.cpp:24-40 (speeds), :57-130 (fields),
    :133-169 (evolved fields from characteristic fields)
  PointwiseFunctions/GeneralRelativity/InterfaceNullNormal.cpp:31-60,
    ProjectionOperators.cpp:24-120
Pinned against fixtures made by importing the reference's Bjorhus.py /
TestFunctions.py (tests/golden/gen_bjorhus_golden.py).

Index conventions: a, b, ... = 0..3 spacetime; i, j, ... = 0..2 spatial, and a
spatial index used in a spacetime slot means a = i + 1.  psi^{ab} = inverse
spacetime metric, g^{ij} = inverse spatial metric, t^a / t_a = unit normal to
the slice, n_i / n^i = outward unit normal of the boundary face.
"""
import numpy as np

_EPS = np.zeros((3, 3, 3))
_EPS[0, 1, 2] = _EPS[1, 2, 0] = _EPS[2, 0, 1] = 1.0
_EPS[0, 2, 1] = _EPS[2, 1, 0] = _EPS[1, 0, 2] = -1.0


def _mixed_spatial_projector(t_up, t_lo):
    """g_a^i = delta_a^i + t^i t_a, shape [a, i]."""
    g = np.zeros((4, 3))
    for i in range(3):
        g[i + 1, i] = 1.0
    return g + np.outer(t_lo, t_up[1:])


def two_index_constraint(d_gauge, t_lo, t_up, ig, ipsi, pi, phi, d_pi, d_phi, gamma2, c3):
    """C_ia, Eq. (44).  d_gauge[a][b] = d_a H_b; d_pi[i][a][b]; d_phi[i][j][a][b] =
    d_i Phi_jab; c3 = three-index constraint C_iab."""
    gm = _mixed_spatial_projector(t_up, t_lo)
    phis = phi[:, 1:, :]                      # Phi_{i k a} with a spatial second index
    tr_dphi = np.einsum("cd,jicd->ji", ipsi, d_phi)        # psi^{cd} d_j Phi_{icd}
    tr_dpi = np.einsum("cd,icd->i", ipsi, d_pi)
    phi_up = np.einsum("ce,df,ief->icd", ipsi, ipsi, phi)  # Phi_i^{cd}
    tr_phi = np.einsum("cd,jcd->j", ipsi, phi)
    c = np.einsum("jk,jika->ia", ig, d_phi[:, :, 1:, :])
    c -= 0.5 * np.einsum("aj,ji->ia", gm, tr_dphi)
    c += np.einsum("b,iba->ia", t_up, d_pi)
    c -= 0.5 * np.outer(tr_dpi, t_lo)
    c += d_gauge[1:, :]
    c += 0.5 * np.einsum("aj,jcd,icd->ia", gm, phi, phi_up)
    c += 0.5 * np.einsum("jk,j,ike,e,a->ia", ig, tr_phi, phis, t_up, t_lo)
    c -= np.einsum("jk,mn,jma,ikn->ia", ig, ig, phis, phis[:, :, 1:])
    tt = np.outer(t_up, t_up)
    c += 0.5 * np.einsum("icd,be,a,cdbe->ia", phi, pi, t_lo,
                         np.einsum("cb,de->cdbe", ipsi, ipsi)
                         + 0.5 * np.einsum("be,cd->cdbe", ipsi, tt))
    c -= np.einsum("icd,ba,c,bd->ia", phi, pi, t_up, ipsi + 0.5 * tt)
    c += 0.5 * gamma2 * np.outer(np.einsum("cd,icd->i", ipsi, c3), t_lo)
    c -= gamma2 * np.einsum("d,iad->ia", t_up, c3)
    return c


def f_constraint(gauge, d_gauge, t_lo, t_up, ig, ipsi, pi, phi, d_pi, d_phi, gamma2, c3):
    """F_a, Eq. (43), without the stress-energy term."""
    gm = _mixed_spatial_projector(t_up, t_lo)
    H_sp, dH_sp = gauge[1:], d_gauge[1:, :]             # H_i, d_i H_b
    pis = pi[1:, :]                                     # Pi_{j a}
    phis = phi[:, 1:, :]                                # Phi_{i j a}
    f = 0.5 * np.einsum("ai,bc,ibc->a", gm, ipsi, d_pi)
    f -= np.einsum("ij,ija->a", ig, d_pi[:, 1:, :])
    f -= np.einsum("ij,b,ijba->a", ig, t_up, d_phi)
    f += 0.5 * t_lo * np.einsum("bc,ij,ijbc->", ipsi, ig, d_phi)
    f += t_lo * np.einsum("ij,ij->", ig, dH_sp[:, 1:])
    f += np.einsum("ai,ijb,jk,kcd,bd,c->a", gm, phis, ig, phi, ipsi, t_up)
    f -= 0.5 * np.einsum("ai,ijb,jk,kcd,cd,b->a", gm, phis, ig, phi, ipsi, t_up)
    f -= np.einsum("ai,b,ib->a", gm, t_up, dH_sp)
    f += np.einsum("ij,icd,jba,bc,d->a", ig, phi, phi, ipsi, t_up)
    f -= 0.5 * t_lo * np.einsum("ij,mn,imc,njd,cd->", ig, ig, phis, phis, ipsi)
    f -= 0.25 * t_lo * np.einsum("ij,icd,jbe,cb,de->", ig, phi, phi, ipsi, ipsi)
    f += 0.25 * t_lo * np.einsum("cd,be,cb,de->", pi, pi, ipsi, ipsi)
    f -= np.einsum("ij,i,ja->a", ig, H_sp, pis)
    f -= np.einsum("b,ij,bi,ja->a", t_up, ig, pi[:, 1:], pis)
    f -= 0.25 * np.einsum("ai,icd,c,d,be,be->a", gm, phi, t_up, t_up, pi, ipsi)
    f += 0.5 * t_lo * np.einsum("cd,be,ce,d,b->", pi, pi, ipsi, t_up, t_up)
    f += np.einsum("ai,icd,be,c,b,de->a", gm, phi, pi, t_up, t_up, ipsi)
    f -= np.einsum("ij,iba,b,je,e->a", ig, phi, t_up, pis, t_up)
    f -= 0.5 * np.einsum("ij,icd,c,d,ja->a", ig, phi, t_up, t_up, pis)
    f -= np.einsum("ij,i,jba,b->a", ig, H_sp, phi, t_up)
    f += np.einsum("ai,icd,b,bc,d->a", gm, phi, gauge, ipsi, t_up)
    # g^{id} C_ida with the spacetime index d running over all four values:
    # g^{id} = psi^{id} + t^i t^d (Constraints.cpp:802-836)
    g_id = ipsi[1:, :] + np.outer(t_up[1:], t_up)
    f += gamma2 * (np.einsum("id,ida->a", g_id, c3)
                   - 0.5 * np.einsum("ai,cd,icd->a", gm, ipsi, c3))
    f += 0.5 * t_lo * np.einsum("cd,cd,b,b->", pi, ipsi, gauge, t_up)
    f -= t_lo * np.einsum("ij,ijc,d,cd->", ig, phis, gauge, ipsi)
    f += 0.5 * t_lo * np.einsum("ij,i,jcd,cd->", ig, H_sp, phi, ipsi)
    return f


def characteristic_speeds(gamma1, lapse, shift, n_lo):
    sdn = float(np.dot(shift, n_lo))
    return np.array([-(1.0 + gamma1) * sdn, -sdn, -sdn + lapse, -sdn - lapse])


def characteristic_fields(gamma2, ig, g, pi, phi, n_lo):
    """(v_psi, v_zero, v_plus, v_minus) of the given (time derivatives of the)
    evolved fields, Characteristics.cpp:57-130."""
    n_up = ig @ n_lo
    phi_n = np.einsum("i,iab->ab", n_up, phi)
    v_zero = phi - np.einsum("i,ab->iab", n_lo, phi_n)
    return g, v_zero, pi + phi_n - gamma2 * g, pi - phi_n - gamma2 * g


def evolved_fields_from_characteristic_fields(gamma2, v_psi, v_zero, v_plus, v_minus, n_lo):
    """Characteristics.cpp:133-169 -> (g, Pi, Phi)."""
    return (v_psi, 0.5 * (v_plus + v_minus) + gamma2 * v_psi,
            np.einsum("i,ab->iab", n_lo, 0.5 * (v_plus - v_minus)) + v_zero)


def physical_terms(gamma2, n_lo, n_up, t_up, p_lo, p_mix, p_up, ig, g, ipsi, c3, rhs_minus, pi,
                   phi, d_phi, d_pi, speeds):
    """add_physical_terms_to_dt_v_minus (BjorhusImpl.cpp:223-494, mu_phys = 0,
    adjust_phys_using_c4, gamma2_in_phys): the incoming Weyl propagating mode
    U^{3-} (WeylPropagating.cpp) from the spatial Ricci tensor of the GH variables
    (GeneralizedHarmonic/Ricci.cpp), the extrinsic curvature (ExtrinsicCurvature.cpp)
    and its covariant derivative (CovariantDerivOfExtrinsicCurvature.cpp)."""
    phis = phi[:, 1:, :]                                 # Phi_{i j a}
    dg3 = phi[:, 1:, 1:]                                 # d_k g_ij
    K = 0.5 * pi[1:, 1:] + 0.5 * (np.einsum("ija,a->ij", phis, t_up)
                                  + np.einsum("jia,a->ij", phis, t_up))
    # Christoffel symbols of the spatial metric: Gamma_{c ab} = (d_b g_ca + d_a g_cb - d_c g_ab)/2
    chr1 = 0.5 * (np.einsum("bca->cab", dg3) + np.einsum("acb->cab", dg3) - dg3)
    chr2 = np.einsum("ij,jkl->ikl", ig, chr1)
    # covariant derivative of K_ij
    w = np.einsum("ca,kcb,b->ka", ipsi, phi, t_up) \
        + 0.5 * np.outer(np.einsum("kcb,c,b->k", phi, t_up, t_up), t_up)
    dphis = d_phi[:, :, 1:, :]                           # d_k Phi_{i j a}
    cdk = (d_pi[:, 1:, 1:] + np.einsum("kija,a->kij", dphis, t_up)
           + np.einsum("kjia,a->kij", dphis, t_up)
           - np.einsum("ija,ka->kij", phis, w) - np.einsum("jia,ka->kij", phis, w))
    cov_dK = 0.5 * cdk - np.einsum("lik,lj->kij", chr2, K) - np.einsum("ljk,li->kij", chr2, K)
    # spatial Ricci tensor from Phi and d Phi
    d3 = d_phi[:, :, 1:, 1:]                             # d_k Phi_{l i j}
    ricci = 0.25 * (np.einsum("kl,jlki->ij", ig, d3) + np.einsum("kl,ilkj->ij", ig, d3)
                    - np.einsum("kl,jikl->ij", ig, d3) - np.einsum("kl,ijkl->ij", ig, d3)
                    + np.einsum("kl,kijl->ij", ig, d3) + np.einsum("kl,kjil->ij", ig, d3)
                    - 2.0 * np.einsum("kl,lkij->ij", ig, d3))
    phi_Ijj = 0.5 * np.einsum("kl,lij->kij", ig, dg3)
    phi_ijK = 0.5 * np.einsum("kl,ijl->ijk", ig, dg3)
    dm2b = np.einsum("kl,lii->k", ig, phi_ijK) - 2.0 * np.einsum("kl,iil->k", ig, phi_Ijj)
    ricci = ricci + 0.5 * (np.einsum("ijk,k->ij", dg3, dm2b) + np.einsum("jik,k->ij", dg3, dm2b)
                           - np.einsum("kij,k->ij", dg3, dm2b)) \
        + np.einsum("ikl,jlk->ij", phi_ijK, phi_ijK) \
        + 2.0 * np.einsum("kil,kjl->ij", phi_Ijj, phi_ijK) \
        - 2.0 * np.einsum("kli,lkj->ij", phi_Ijj, phi_Ijj)
    # adjust_phys_using_c4: add multiples of the four-index constraint
    ricci = ricci + 0.25 * (np.einsum("kl,iklj->ij", ig, d3) - np.einsum("kl,kilj->ij", ig, d3)
                            + np.einsum("kl,jkli->ij", ig, d3) - np.einsum("kl,kjli->ij", ig, d3))
    ricci = ricci + 0.5 * (np.einsum("k,a,ikja->ij", n_up, t_up, dphis)
                           - np.einsum("k,a,kija->ij", n_up, t_up, dphis)
                           + np.einsum("k,a,jkia->ij", n_up, t_up, dphis)
                           - np.einsum("k,a,kjia->ij", n_up, t_up, dphis))
    # Weyl electric part and the incoming propagating mode (sign = -1)
    weyl_e = ricci + np.einsum("kl,kl->", K, ig) * K - np.einsum("il,kl,kj->ij", K, ig, K)
    sign = -1.0
    tmp = weyl_e - sign * np.einsum("k,kij->ij", n_up, cov_dK) \
        + sign * 0.5 * np.einsum("k,jik->ij", n_up, cov_dK) \
        + sign * 0.5 * np.einsum("k,ijk->ij", n_up, cov_dK)
    sp_up = ig - np.outer(n_up, n_up)
    sp_lo = g[1:, 1:] - np.outer(n_lo, n_lo)
    sp_mix = np.eye(3) - np.outer(n_up, n_lo)
    weyl_minus = np.einsum("ki,lj,kl->ij", sp_mix, sp_mix, tmp) \
        - 0.5 * np.einsum("kl,kl->", sp_up, tmp) * sp_lo
    u3_minus = 2.0 * np.einsum("ia,jb,ij->ab", p_mix[1:, :], p_mix[1:, :], weyl_minus)
    t = speeds[3] * u3_minus - speeds[3] * gamma2 * np.einsum("i,iab->ab", n_up, c3)
    total = rhs_minus + t
    return np.einsum("ac,bd,ab->cd", p_mix, p_mix, total) \
        - 0.5 * np.einsum("ab,ab->", p_up, total) * p_lo


def bjorhus_constraint_preserving(n_lo, g, pi, phi, coords, gamma1, gamma2, lapse, shift, ipsi,
                                  t_up, c3, gauge, d_gauge, dt_g, dt_pi, dt_phi, d_pi, d_phi,
                                  physical=False):
    """ConstraintPreservingBjorhus::dg_time_derivative for Type ConstraintPreserving
    (physical=False) or ConstraintPreservingPhysical (physical=True) on a static
    mesh, at one face point.  Arguments as the reference passes them
    (Bjorhus.cpp:104-148); returns the corrections (dt g, dt Pi, dt Phi) that are
    ADDED to the volume time derivative on the boundary points."""
    t_lo = np.zeros(4)
    t_lo[0] = -lapse
    ig = ipsi[1:, 1:] + np.outer(shift, shift) / (lapse * lapse)
    n_up = ig @ n_lo
    # compute_intermediate_vars
    c4 = np.einsum("ijk,jkab->iab", _EPS, d_phi)
    r2 = np.sqrt(0.5)
    n4_lo, n4_up = np.concatenate([[0.0], n_lo]), np.concatenate([[0.0], n_up])
    in_lo, out_lo = r2 * (t_lo - n4_lo), r2 * (t_lo + n4_lo)
    in_up, out_up = r2 * (t_up - n4_up), r2 * (t_up + n4_up)
    p_lo = g + np.outer(t_lo, t_lo) - np.outer(n4_lo, n4_lo)
    p_mix = np.eye(4) + np.outer(t_up, t_lo) - np.outer(n4_up, n4_lo)      # P^a_b
    p_up = ipsi + np.outer(t_up, t_up) - np.outer(n4_up, n4_up)
    rhs_psi, rhs_zero, rhs_plus, rhs_minus = characteristic_fields(gamma2, ig, dt_g, dt_pi,
                                                                   dt_phi, n_lo)
    c2 = two_index_constraint(d_gauge, t_lo, t_up, ig, ipsi, pi, phi, d_pi, d_phi, gamma2, c3)
    fc = f_constraint(gauge, d_gauge, t_lo, t_up, ig, ipsi, pi, phi, d_pi, d_phi, gamma2, c3)
    nc2 = np.einsum("i,ia->a", n_up, c2)
    c0_plus, c0_minus = fc - nc2, fc + nc2
    speeds = characteristic_speeds(gamma1, lapse, shift, n_lo)
    if speeds.min() >= 0.0:
        return np.zeros((4, 4)), np.zeros((4, 4)), np.zeros((3, 4, 4))
    # BjorhusImpl.cpp:26-47: dt v_psi = lambda_psi n^i C_iab
    bc_psi = speeds[0] * np.einsum("i,iab->ab", n_up, c3)
    # :49-103: dt v_zero_iab = lambda_0 n^k eps_{i j k}-contracted four-index constraint
    bc_zero = np.zeros((3, 4, 4))
    for i in range(3):
        for j in range(3):
            for k in range(3):
                if _EPS[i, j, k] != 0.0:
                    bc_zero[i] += _EPS[i, j, k] * speeds[1] * n_up[k] * c4[j]
    bc_plus = -rhs_plus
    # :153-221 constraint-dependent terms (mu = 0)
    A = rhs_minus
    t1 = np.einsum("c,d,cd->", in_up, in_up, A) * np.outer(out_lo, out_lo)
    in_A = np.einsum("c,cd->d", in_up, A)           # u^c A_cd
    A_in = np.einsum("d,cd->c", in_up, A)           # A_cd u^d
    t2 = np.outer(np.einsum("da,d->a", p_mix, in_A), out_lo)
    t3 = np.outer(out_lo, np.einsum("db,d->b", p_mix, in_A))
    t4 = np.outer(np.einsum("ca,c->a", p_mix, A_in), out_lo)
    t5 = np.outer(out_lo, np.einsum("cb,c->b", p_mix, A_in))
    t6 = np.einsum("cd,cd->", p_up, A) * p_lo
    common = r2 * speeds[3] * c0_minus
    t7 = np.dot(in_up, common) * np.outer(out_lo, out_lo)
    t8 = np.dot(out_up, common) * p_lo
    pc = np.einsum("ca,c->a", p_mix, common)
    t9 = np.outer(out_lo, pc)
    t10 = np.outer(pc, out_lo)
    bc_minus = 0.5 * (2.0 * t1 - t2 - t3 - t4 - t5 + t6) + (t7 + t8 - t9 - t10)
    # :105-151 gauge Sommerfeld terms
    prefac = gamma2 - 1.0 / np.sqrt(np.sum(np.asarray(coords) ** 2))
    B = rhs_psi
    pB = np.einsum("cb,d,cd->b", p_mix, out_up, B)
    s1 = np.outer(in_lo, pB)
    s2 = np.outer(pB, in_lo)
    uBv = np.einsum("c,d,cd->", in_up, out_up, B)
    s3 = uBv * np.outer(in_lo, out_lo)
    s4 = uBv * np.outer(out_lo, in_lo)
    s5 = np.einsum("c,d,cd->", out_up, out_up, B) * np.outer(in_lo, in_lo)
    bc_minus = bc_minus + prefac * (s1 + s2 - s3 - s4 - s5) - rhs_minus
    if physical:  # BjorhusImpl.cpp:534-593
        bc_minus = bc_minus + physical_terms(gamma2, n_lo, n_up, t_up, p_lo, p_mix, p_up, ig, g,
                                             ipsi, c3, rhs_minus, pi, phi, d_phi, d_pi, speeds)
    # only incoming characteristic fields are corrected (Bjorhus.cpp:38-47, :345-352)
    if speeds[0] > 0.0:
        bc_psi = np.zeros_like(bc_psi)
    if speeds[1] > 0.0:
        bc_zero = np.zeros_like(bc_zero)
    if speeds[2] > 0.0:
        bc_plus = np.zeros_like(bc_plus)
    if speeds[3] > 0.0:
        bc_minus = np.zeros_like(bc_minus)
    return evolved_fields_from_characteristic_fields(gamma2, bc_psi, bc_zero, bc_plus, bc_minus,
                                                     n_lo)
