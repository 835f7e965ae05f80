// pack.cpp -- flatten the host data model (ragged Atom->Shell->array objects of the reference,
// Objects.f90:184-275) into the SoA / CSR image declared in include/trekis3_gpu.h.
#include "trk3_host.hpp"
#include <cstring>

namespace trk3 {

namespace {
void pack_rows(const DiffCS &d, std::vector<int64_t> &off, std::vector<double> &hw, std::vector<double> &L) {
    for (auto &r : d.row) {
        hw.insert(hw.end(), r.hw.begin(), r.hw.end());
        L.insert(L.end(), r.L.begin(), r.L.end());
        off.push_back((int64_t)hw.size());
    }
}
}  // namespace

void pack_case(const Case &c, Packed &p) {
    p = Packed{};
    trk3_config &g = p.cfg;
    g.shi_E = c.SHI.E; g.shi_mass = c.SHI.Mass; g.shi_fixed_Zeff = c.SHI.fixed_Zeff; g.shi_Z = c.SHI.Zat; g.shi_kind_Zeff = c.SHI.Kind_Zeff;
    g.Tim = c.Tim; g.dt = c.dt; g.dt_flag = c.numpar.dt_flag; g.include_photons = c.numpar.include_photons ? 1 : 0;
    g.cut_off = c.Matter.cut_off; g.layer = c.Matter.Layer; g.hole_mass = c.Matter.hole_mass;
    g.work_function = c.Matter.work_function; g.bar_length = c.Matter.bar_length; g.bar_height = c.Matter.bar_height;
    g.kind_of_EMFP = c.numpar.kind_of_EMFP; g.seed = 20260101ull;

    trk3_tables &t = p.tab;
    const int Nat = (int)c.atoms.size();
    t.n_atoms = Nat; t.n_shells = c.n_shells(); t.nshl_atom1 = c.atoms[0].nshl();
    int s = 0;
    for (int a = 0; a < Nat; ++a) {
        const Atom &at = c.atoms[a];
        t.atom_Z[a] = at.Zat; t.atom_nshl[a] = at.nshl(); t.atom_first[a] = s; t.atom_mass[a] = at.Mass; t.atom_pers[a] = at.Pers;
        for (int k = 0; k < at.nshl(); ++k, ++s) {
            t.shell_atom[s] = a; t.shell_num[s] = k; t.shell_Ip[s] = at.Ip[k]; t.shell_Nel[s] = at.Nel[k];
            t.shell_auger[s] = at.Auger[k]; t.shell_radiat[s] = at.Radiat[k];
            t.shell_kocs[s] = (k < (int)at.KOCS.size() && at.KOCS[k] == 2) ? 2 : 1; t.shell_Ek[s] = k < (int)at.Ek.size() ? at.Ek[k] : 0.0;
            if (a == c.Lowest_Ip_At && k == c.Lowest_Ip_Shl) t.vb_shell = s;
        }
    }
    t.at_dens = c.Matter.At_Dens;
    // oscillators of the delta-function CDF per flat shell (kind_of_DR = 4)
    t.delta_cdf = (c.numpar.kind_of_DR == 4) ? 1 : 0;
    p.osc_E0.clear(); p.osc_alpha.clear();
    {
        int sf = 0;
        for (int a = 0; a < t.n_atoms; ++a) for (int k = 0; k < c.atoms[a].nshl(); ++k, ++sf) {
            t.osc_off[sf] = (int32_t)p.osc_E0.size();
            const CDFosc &o = c.atoms[a].Ritchi[k];
            if (t.delta_cdf) for (size_t l = 0; l < o.E0.size() && l < o.alpha.size(); ++l) { p.osc_E0.push_back(o.E0[l]); p.osc_alpha.push_back(o.alpha[l]); }
        }
        for (; sf <= TRK3_MAX_SHELLS; ++sf) t.osc_off[sf] = (int32_t)p.osc_E0.size();
    }
    if (p.osc_E0.empty()) { p.osc_E0.push_back(0.0); p.osc_alpha.push_back(0.0); }
    t.osc_E0 = p.osc_E0.data(); t.osc_alpha = p.osc_alpha.data();
    // DSF elastic scattering (kind_of_EMFP = 2): rows of the two DSF files, flattened [particle energy][transferred energy]
    auto pack_dsf = [](const std::vector<DsfPoint> &D, int32_t &n, std::vector<double> &dE, std::vector<double> &em, std::vector<double> &ab,
                       std::vector<double> &Lem, std::vector<double> &Lab) {
        dE.clear(); em.clear(); ab.clear(); Lem.clear(); Lab.clear();
        n = D.empty() ? 0 : (int32_t)D[0].dE.size();
        for (const DsfPoint &d : D) {
            dE.insert(dE.end(), d.dE.begin(), d.dE.end()); em.insert(em.end(), d.dL_emit.begin(), d.dL_emit.end()); ab.insert(ab.end(), d.dL_absorb.begin(), d.dL_absorb.end());
            Lem.push_back(d.dL_emit.back()); Lab.push_back(d.dL_absorb.back());
        }
        if (dE.empty()) { dE.push_back(0.0); em.push_back(0.0); ab.push_back(0.0); Lem.push_back(0.0); Lab.push_back(0.0); }
    };
    const bool dsf = c.numpar.kind_of_EMFP == 2;
    static const std::vector<DsfPoint> none;
    pack_dsf(dsf ? c.DSF_DEMFP : none, t.n_dsf_e, p.dsf_e_dE, p.dsf_e_emit, p.dsf_e_absorb, p.ee_emit, p.ee_absorb);
    pack_dsf(dsf ? c.DSF_DEMFP_H : none, t.n_dsf_h, p.dsf_h_dE, p.dsf_h_emit, p.dsf_h_absorb, p.he_emit, p.he_absorb);
    t.dsf_e_dE = p.dsf_e_dE.data(); t.dsf_e_emit = p.dsf_e_emit.data(); t.dsf_e_absorb = p.dsf_e_absorb.data(); t.ee_emit = p.ee_emit.data(); t.ee_absorb = p.ee_absorb.data();
    t.dsf_h_dE = p.dsf_h_dE.data(); t.dsf_h_emit = p.dsf_h_emit.data(); t.dsf_h_absorb = p.dsf_h_absorb.data(); t.he_emit = p.he_emit.data(); t.he_absorb = p.he_absorb.data();
    const int NS = t.n_shells;
    auto pack_mfp = [&](const std::vector<std::vector<MFP>> &T, std::vector<double> &E, std::vector<double> &L, std::vector<double> *dEdx) {
        E = T[0][0].E;
        const size_t N = E.size();
        L.assign((size_t)NS * N, 0.0);
        if (dEdx) dEdx->assign((size_t)NS * N, 0.0);
        int q = 0;
        for (int a = 0; a < Nat; ++a) for (int k = 0; k < c.atoms[a].nshl(); ++k, ++q) {
            std::memcpy(&L[(size_t)q * N], T[a][k].L.data(), N * 8);
            if (dEdx) std::memcpy(&(*dEdx)[(size_t)q * N], T[a][k].dEdx.data(), N * 8);
        }
    };
    pack_mfp(c.Total_el_MFPs, p.ei_E, p.ei_L, nullptr);
    pack_mfp(c.Total_Hole_MFPs, p.hi_E, p.hi_L, nullptr);
    pack_mfp(c.SHI_MFP, p.shi_E, p.shi_L, &p.shi_dEdx);
    if (!c.Total_Photon_MFPs.empty()) pack_mfp(c.Total_Photon_MFPs, p.ph_E, p.ph_L, nullptr);
    p.ee_E = c.Elastic_MFP.E; p.ee_L = c.Elastic_MFP.L;
    p.he_E = c.Elastic_Hole_MFP.E; p.he_L = c.Elastic_Hole_MFP.L;
    // differential tables
    p.dshi_off.push_back(0);
    for (int a = 0; a < Nat; ++a) for (int k = 0; k < c.atoms[a].nshl(); ++k) {
        const MFP &m = c.diff_SHI_MFP[a][k];
        p.dshi_E.insert(p.dshi_E.end(), m.E.begin(), m.E.end());
        p.dshi_L.insert(p.dshi_L.end(), m.L.begin(), m.L.end());
        p.dshi_off.push_back((int64_t)p.dshi_E.size());
    }
    p.eid_off.push_back(0);
    for (int a = 0; a < Nat; ++a) for (int k = 0; k < c.atoms[a].nshl(); ++k) pack_rows(c.EIdCS[a][k], p.eid_off, p.eid_hw, p.eid_L);
    p.eed_off.push_back(0); pack_rows(c.EEdCS, p.eed_off, p.eed_hw, p.eed_L);
    p.hid_off.push_back(0); pack_rows(c.HIdCS, p.hid_off, p.hid_hw, p.hid_L);
    p.hed_off.push_back(0); pack_rows(c.HEdCS, p.hed_off, p.hed_hw, p.hed_L);
    p.dos_E = c.dos.E; p.dos_DOS = c.dos.dos; p.dos_int = c.dos.int_DOS; p.dos_effm = c.dos.Eff_m;
    p.out_R = c.Out_R; p.out_V = c.Out_V;

    t.n_ei = (int)p.ei_E.size(); t.ei_E = p.ei_E.data(); t.ei_L = p.ei_L.data();
    t.n_ee = (int)p.ee_E.size(); t.ee_E = p.ee_E.data(); t.ee_L = p.ee_L.data();
    t.n_hi = (int)p.hi_E.size(); t.hi_E = p.hi_E.data(); t.hi_L = p.hi_L.data();
    t.n_he = (int)p.he_E.size(); t.he_E = p.he_E.data(); t.he_L = p.he_L.data();
    t.n_ph = (int)p.ph_E.size(); t.ph_E = p.ph_E.empty() ? nullptr : p.ph_E.data(); t.ph_L = p.ph_L.empty() ? nullptr : p.ph_L.data();
    t.n_shi = (int)p.shi_E.size(); t.shi_E = p.shi_E.data(); t.shi_L = p.shi_L.data(); t.shi_dEdx = p.shi_dEdx.data();
    t.dshi_off = p.dshi_off.data(); t.dshi_E = p.dshi_E.data(); t.dshi_L = p.dshi_L.data();
    t.eid_off = p.eid_off.data(); t.eid_hw = p.eid_hw.data(); t.eid_L = p.eid_L.data();
    t.eed_off = p.eed_off.data(); t.eed_hw = p.eed_hw.data(); t.eed_L = p.eed_L.data();
    t.hid_off = p.hid_off.data(); t.hid_hw = p.hid_hw.data(); t.hid_L = p.hid_L.data();
    t.hed_off = p.hed_off.data(); t.hed_hw = p.hed_hw.data(); t.hed_L = p.hed_L.data();
    t.n_dos = (int)p.dos_E.size(); t.dos_E = p.dos_E.data(); t.dos_DOS = p.dos_DOS.data(); t.dos_int = p.dos_int.data(); t.dos_effm = p.dos_effm.data();
    t.n_r = (int)p.out_R.size(); t.out_R = p.out_R.data(); t.out_V = p.out_V.data();
}

}  // namespace trk3
