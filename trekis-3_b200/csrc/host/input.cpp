// input.cpp -- readers of the reference's input contract:
//   INPUT_PARAMETERS.txt            Reading_files_and_parameters.f90:162-634, 948-1116
//   INPUT_CDF/<material>.cdf        Reading_files_and_parameters.f90:1185-1644
//   INPUT_DOS/<material>.dos        Reading_files_and_parameters.f90:2317-2454
//   INPUT_EADL/INPUT_atomic_data.dat (periodic table; same numbers as Dealing_with_EADL.f90:1217-1705)
// The files are parsed with the semantics of Fortran list-directed input (blank lines are
// skipped, the rest of a record after the requested items is ignored, d-exponents are legal).
#include "trk3_host.hpp"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <sys/stat.h>

namespace trk3 {

namespace {

struct Lines {
    std::vector<std::string> l;
    size_t pos = 0;
    bool load(const std::string &path) {
        std::ifstream f(path);
        if (!f) return false;
        std::string s;
        while (std::getline(f, s)) {
            if (!s.empty() && s.back() == '\r') s.pop_back();
            l.push_back(s);
        }
        return true;
    }
    bool eof() const { return pos >= l.size(); }
    bool read_line(std::string &s) { if (eof()) return false; s = l[pos++]; return true; }
    void backspace() { if (pos > 0) --pos; }
};

// split a record into list-directed tokens (blank, tab, comma separated; quotes honoured)
std::vector<std::string> tokens(const std::string &s) {
    std::vector<std::string> t;
    size_t i = 0, n = s.size();
    while (i < n) {
        while (i < n && (s[i] == ' ' || s[i] == '\t' || s[i] == ',')) ++i;
        if (i >= n) break;
        if (s[i] == '\'' || s[i] == '"') {
            char q = s[i++]; size_t j = i;
            while (j < n && s[j] != q) ++j;
            t.push_back(s.substr(i, j - i)); i = (j < n) ? j + 1 : n;
        } else {
            size_t j = i;
            while (j < n && s[j] != ' ' && s[j] != '\t' && s[j] != ',') ++j;
            t.push_back(s.substr(i, j - i)); i = j;
        }
    }
    return t;
}

bool parse_real(const std::string &tok, double &v) {
    if (tok.empty()) return false;
    std::string s = tok;
    for (auto &ch : s) if (ch == 'd' || ch == 'D') ch = 'e';
    char *end = nullptr;
    v = std::strtod(s.c_str(), &end);
    return end && *end == '\0' && end != s.c_str();
}
bool parse_int(const std::string &tok, int &v) {
    if (tok.empty()) return false;
    char *end = nullptr;
    long x = std::strtol(tok.c_str(), &end, 10);
    if (!(end && *end == '\0' && end != tok.c_str())) return false;   // "1.0" is not a valid integer item
    v = (int)x; return true;
}

// READ(unit,*) of n items: consume records until n tokens are collected.
// reason: 0 ok, -1 end of file, +1 conversion error (checked by the caller through the parse_* calls)
int read_list(Lines &f, size_t n, std::vector<std::string> &out) {
    out.clear();
    while (out.size() < n) {
        std::string s;
        if (!f.read_line(s)) return -1;
        auto t = tokens(s);
        for (auto &x : t) { if (out.size() < n) out.push_back(x); }
    }
    return 0;
}

bool file_exists(const std::string &p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }

struct PT { std::string name, full; double mass, nvb; };
bool load_periodic_table(const std::string &dir, std::vector<PT> &pt, std::string &err) {
    std::string path = dir + "/INPUT_EADL/INPUT_atomic_data.dat";
    Lines f;
    if (!f.load(path)) { err = "File " + path + " is not found!"; return false; }
    pt.assign(130, PT{"", "", 0.0, 0.0});
    for (size_t i = 1; i < f.l.size(); ++i) {
        auto t = tokens(f.l[i]);
        int z; double m, nvb;
        if (t.size() < 5 || !parse_int(t[0], z) || !parse_real(t[3], m) || !parse_real(t[4], nvb)) continue;
        if (z > 0 && z < 130) pt[z] = PT{t[2], t[1], m, nvb};
    }
    // The reference takes element names/masses from the hard-coded Find_element_name
    // (Dealing_with_EADL.f90:1217-1705), which differs from the data file for these entries:
    auto fix = [&](int z, const char *n, const char *fn, double m) { if (z < (int)pt.size()) { pt[z].name = n; pt[z].full = fn; pt[z].mass = m; } };
    fix(56, "Ba", "Barium", 137.327); fix(77, "Ir", "Iridium", 192.217); fix(94, "Pu", "Plutonium", 244.0);
    fix(111, "Rg", "Roentgenium", 281.0); fix(115, "Uup", "Ununpentium", 288.0);
    fix(117, "Uus", "Ununseptium", 294.0); fix(118, "Uuo", "Ununoctium", 294.0);
    return true;
}

// define_PQN, Dealing_with_EADL.f90:1016-1214: ENDL shell designator -> name, principal quantum number
void define_PQN(int d, std::string &name, int &pqn) {
    static const char *names[] = {"", "K-shell", "L-shell", "L1-shell", "L23-shell", "L2-shell", "L3-shell", "M-shell",
        "M1-shell", "M23-shell", "M2-shell", "M3-shell", "M45-shell", "M4-shell", "M5-shell", "N-shell", "N1-shell",
        "N23-shell", "N2-shell", "N3-shell", "N45-shell", "N4-shell", "N5-shell", "N67-shell", "N6-shell", "N7-shell",
        "O-shell", "O1-shell", "O23-shell", "O2-shell", "O3-shell", "O45-shell", "O4-shell", "O5-shell", "O67-shell",
        "O6-shell", "O7-shell", "O89-shell", "O8-shell", "O9-shell", "P-shell", "P1-shell", "P23-shell", "P2-shell",
        "P3-shell", "P45-shell", "P4-shell", "P5-shell", "P67-shell", "P6-shell", "P7-shell", "P89-shell", "P8-shell",
        "P9-shell", "P1011-shell", "P10-shell", "P11-shell", "Q-shell", "Q1-shell", "Q23-shell", "Q2-shell", "Q3-shell"};
    if (d >= 63) { name = "Valence"; pqn = 0; return; }
    if (d >= 1 && d <= 61) name = names[d]; else name = "Shell";
    if (d == 1) pqn = 1; else if (d <= 6) pqn = 2; else if (d <= 14) pqn = 3; else if (d <= 25) pqn = 4;
    else if (d <= 39) pqn = 5; else if (d <= 56) pqn = 6; else pqn = 7;
}

// interpret_additional_data_INPUT, Reading_files_and_parameters.f90:948-1116
void interpret_flag(const std::string &line, NumPar &np) {
    auto t = tokens(line);
    if (t.empty()) return;
    const std::string &k = t[0];
    auto is = [&](std::initializer_list<const char *> names) { for (auto n : names) if (k == n) return true; return false; };
    int iv;
    if (is({"grid", "GRID", "Grid"})) { if (t.size() > 1 && parse_int(t[1], iv)) np.CS_method = iv; }
    else if (is({"UNITS", "Units", "units"})) { if (t.size() > 1 && parse_int(t[1], iv)) np.out_dim = iv; }
    else if (is({"CDF", "Cdf", "cdf"})) { if (t.size() > 1) np.CDF_file = t[1]; }
    else if (is({"DOS", "Dos", "dos"})) { if (t.size() > 1) np.DOS_file = t[1]; }
    else if (is({"gnuplot", "plot", "gnu", "GNUPLOT", "PLOT", "GNU"})) {
        np.do_gnuplot = true;
        np.plot_extension = (t.size() > 1) ? t[1] : "jpeg";
        if (np.plot_extension == "NO" || np.plot_extension == "No" || np.plot_extension == "no") np.do_gnuplot = false;
    }
    else if (is({"redo_MFP", "REDO_MFP", "Redo_MFP", "redo_mfp"})) { np.redo_IMFP = np.redo_EMFP = np.redo_IMFP_SHI = true; }
    else if (is({"redo_MFP_SHI", "REDO_MFP_SHI", "Redo_MFP_SHI", "redo_mfp_shi", "redo_IMFP_SHI", "REDO_IMFP_SHI", "Redo_IMFP_SHI", "redo_imfp_shi"})) np.redo_IMFP_SHI = true;
    else if (is({"redo_IMFP", "REDO_IMFP", "Redo_IMFP", "redo_imfp"})) np.redo_IMFP = true;
    else if (is({"redo_EMFP", "REDO_EMFP", "Redo_EMFP", "redo_emfp"})) np.redo_EMFP = true;
    else if (is({"get_thermal", "thermal", "make_thermal", "Get_thermal", "Thermal", "Make_thermal"})) np.get_thermal = true;
    else if (is({"print_CDF", "Print_CDF", "print_cdf", "PRINT_CDF"})) np.print_CDF = true;
    else if (is({"print_optical", "Print_optical", "Print_Optical", "print_optical_cdf", "PRINT_OPTICAL_CDF"})) np.print_CDF_optical = true;
    else if (is({"Verbose", "verbose", "VERBOSE"})) np.verbose = true;
    else if (is({"Very_verbose", "very_verbose", "VERY_VERBOSE", "Very_Verbose", "Very-verbose", "very-verbose", "VERY-VERBOSE", "Very-Verbose"})) { np.verbose = true; np.very_verbose = true; }
}

bool read_input_parameters(const std::string &dir, Case &c, const std::vector<PT> &pt, std::string &err) {
    Lines f;
    std::string path = dir + "/INPUT_PARAMETERS.txt";
    if (!f.load(path)) { err = "File INPUT_PARAMETERS.txt is not found!"; return false; }
    c.input_lines = f.l;
    std::vector<std::string> t;
    int line = 0;
    auto fail = [&](const char *what) { err = std::string("Problem reading INPUT_PARAMETERS.txt in line ") + std::to_string(line) + ": " + what; return false; };
    auto need = [&](size_t n) { ++line; return read_list(f, n, t) == 0; };
    double M;
    if (!need(1)) return fail("material name");
    c.Material_name = t[0];
    if (!need(1) || !parse_int(t[0], c.SHI.Zat)) return fail("SHI atomic number");
    if (c.SHI.Zat > 0) {
        int z = std::abs(c.SHI.Zat);
        if (z >= (int)pt.size() || pt[z].mass <= 0) return fail("unknown SHI element");
        c.SHI.Name = pt[z].name; c.SHI.Full_Name = pt[z].full; c.SHI.Mass = pt[z].mass;
    }
    if (!need(1) || !parse_real(t[0], c.SHI.E)) return fail("SHI energy");
    c.SHI.E *= 1.0e6;
    if (!need(1) || !parse_real(t[0], M)) return fail("SHI mass");
    if (M > 0.0) c.SHI.Mass = M;
    if (!need(1) || !parse_real(t[0], c.Tim)) return fail("total time");
    if (!need(2) || !parse_real(t[0], c.dt) || !parse_int(t[1], c.numpar.dt_flag)) return fail("time step");
    if (c.dt > c.Tim && c.numpar.dt_flag <= 0) c.dt = c.Tim;
    if (c.dt < 2 && c.numpar.dt_flag >= 1) c.dt = 2;
    if (!need(1) || !parse_real(t[0], c.Matter.cut_off)) return fail("cut-off");
    if (!need(1) || !parse_real(t[0], c.Matter.Layer)) return fail("layer");
    if (!need(1) || !parse_real(t[0], c.Matter.temp)) return fail("temperature");
    if (!need(2) || !parse_int(t[0], c.SHI.Kind_Zeff) || !parse_real(t[1], c.SHI.fixed_Zeff)) return fail("Zeff kind");
    if (c.SHI.fixed_Zeff <= 0.0 || c.SHI.fixed_Zeff > c.SHI.Zat) c.SHI.fixed_Zeff = c.SHI.Zat;
    if (!need(1) || !parse_int(t[0], c.SHI.Kind_ion)) return fail("kind of ion");
    if (!need(2) || !parse_int(t[0], c.numpar.kind_of_EMFP) || !parse_int(t[1], c.numpar.CDF_elast_Zeff)) return fail("elastic model");
    if (!need(2) || !parse_int(t[0], c.numpar.kind_of_DR) || !parse_real(t[1], c.Matter.El_eff_mass)) return fail("dispersion relation");
    if (c.numpar.kind_of_DR <= 0 || c.numpar.kind_of_DR > 4) c.numpar.kind_of_DR = 1;
    if (c.Matter.El_eff_mass < 0) c.Matter.El_eff_mass = 1.0;
    int tmp;
    if (!need(1) || !parse_int(t[0], tmp)) return fail("plasmon flag");
    c.numpar.plasmon_Emax = (tmp == 1);
    if (!need(1) || !parse_real(t[0], c.Matter.hole_mass)) return fail("hole mass");
    if (!need(1) || !parse_int(t[0], tmp)) return fail("photon flag");
    c.numpar.include_photons = (tmp == 1);
    if (!need(3) || !parse_real(t[0], c.Matter.work_function) || !parse_real(t[1], c.Matter.bar_length) || !parse_real(t[2], c.Matter.bar_height)) return fail("work function");
    if (c.Matter.work_function <= 0.0) c.Matter.work_function = 0.0;
    if (c.Matter.bar_length <= 0.0) c.Matter.work_function = 0.0;
    if (c.Matter.bar_height <= 0.0) c.Matter.work_function = 0.0;
    if (!need(1) || !parse_int(t[0], c.NMC)) return fail("NMC");
    if (!need(1) || !parse_int(t[0], c.Num_th)) return fail("threads");
    std::string s;
    while (f.read_line(s)) interpret_flag(s, c.numpar);
    return true;
}

// Decompose_compound + sort_elements_in_chem_formula, Dealing_with_EADL.f90:28-198: "Al2O3" -> (Al, 2), (O, 3), heaviest first.
// A digit extends the count of the current element, an upper-case letter starts a new element, lower-case letters continue
// its name; any other character is skipped with a warning, as the reference does.
bool decompose_compound(const std::string &formula, const std::vector<PT> &pt, std::vector<Atom> &atoms, std::vector<std::string> &warnings, std::string &err) {
    std::vector<std::pair<std::string, double>> el;
    int C = 0;
    std::string cur;
    auto close = [&]() { if (!cur.empty()) el.push_back({cur, C <= 0 ? 1.0 : (double)C}); };
    for (char ch : formula) {
        if (ch >= '0' && ch <= '9') C = C * 10 + (ch - '0');
        else if (ch >= 'A' && ch <= 'Z') { close(); C = 0; cur = std::string(1, ch); }
        else if (ch >= 'a' && ch <= 'z') cur += ch;
        else if (ch != ' ' && ch != '\t') warnings.push_back(std::string("Symbol ") + ch + " in the compound formula could not be identified");
    }
    close();
    if (el.empty()) { err = "no element found in the chemical formula '" + formula + "'"; return false; }
    atoms.assign(el.size(), Atom{});
    for (size_t i = 0; i < el.size(); ++i) {
        int Z = 0;
        for (size_t z = 1; z < pt.size(); ++z) {
            std::string nm = pt[z].name;
            while (!nm.empty() && nm.back() == ' ') nm.pop_back();
            if (!nm.empty() && nm == el[i].first) { Z = (int)z; break; }
        }
        if (Z == 0) { err = el[i].first + " - such an element was not found in our database..."; return false; }
        atoms[i].Zat = Z; atoms[i].Pers = el[i].second; atoms[i].Name = el[i].first; atoms[i].Full_Name = pt[(size_t)Z].full; atoms[i].Mass = pt[(size_t)Z].mass;
    }
    for (int j = (int)atoms.size() - 1; j >= 1; --j) {       // bubble sort, descending atomic number (:160-193)
        bool swapped = false;
        for (int i = 0; i < j; ++i) if (atoms[(size_t)i].Zat < atoms[(size_t)i + 1].Zat) { std::swap(atoms[(size_t)i], atoms[(size_t)i + 1]); swapped = true; }
        if (!swapped) break;
    }
    return true;
}

// check_atomic_parameters, ALL-shells branch (Dealing_with_EADL.f90:374-392) with READ_EADL_TYPE_FILE_int / _real (:506-535,
// :661-674): every sub-shell EADL lists for the element, in the order of the file.  The real-valued blocks are copied BY
// POSITION (the I = 921 / 922 blocks list fewer sub-shells than I = 912: the outer ones keep the 1e-24 they were allocated with,
// which the conversion below turns into a practically infinite decay time).
bool eadl_all_shells(const Eadl &db, Atom &a, bool include_photons, std::string &err) {
    const Eadl::Block *b = db.find(a.Zat, 912);
    if (!b || b->rows.empty()) { err = "element Z=" + std::to_string(a.Zat) + " is not in the EADL database"; return false; }
    const size_t n = b->rows.size();
    a.Shell_name.assign(n, ""); a.Shl_num.assign(n, 0); a.PQN.assign(n, 0); a.Nel.assign(n, 0.0);
    a.Ip.assign(n, 0.0); a.Ek.assign(n, 0.0); a.Radiat.assign(n, 1.0e-24); a.Auger.assign(n, 1.0e-24);
    a.KOCS.assign(n, 0); a.KOCS_SHI.assign(n, 0); a.Ritchi.assign(n, CDFosc{});
    for (size_t i = 0; i < n; ++i) {
        a.Nel[i] = b->rows[i].val; a.Shl_num[i] = (int)b->rows[i].des;
        define_PQN(a.Shl_num[i], a.Shell_name[i], a.PQN[i]);
    }
    auto fill = [&](int I, std::vector<double> &arr) {
        const Eadl::Block *r = db.find(a.Zat, I);
        if (!r) { std::fill(arr.begin(), arr.end(), 1.0e-30); return; }          // "if the value does not exist"
        for (size_t i = 0; i < r->rows.size() && i < arr.size(); ++i) arr[i] = r->rows[i].val * 1.0e6;
    };
    fill(913, a.Ip); fill(914, a.Ek); fill(921, a.Radiat); fill(922, a.Auger);
    for (size_t j = 0; j < n; ++j) {
        a.Auger[j] = 1.0e15 * g_h / (g_e * a.Auger[j]);
        a.Radiat[j] = (a.Shl_num[j] >= 63 || !include_photons) ? 1.0e23 : 1.0e15 * g_h / (g_e * a.Radiat[j]);
    }
    return true;
}

// make_valence_band + copy_atomic_data_back, Reading_files_and_parameters.f90:1989-2137: the outermost atomic shells that hold the
// element's valence electrons (INPUT_atomic_data.dat, column 5) are merged into ONE valence band, kept as the last shell of the
// first element: N_e_VB electrons per molecule, ionisation potential = band gap.
void make_valence_band(std::vector<Atom> &atoms, const std::vector<PT> &pt, Solid &Matter) {
    std::vector<size_t> n_core(atoms.size(), 0);
    double N_e_VB = 0.0;
    for (size_t i = 0; i < atoms.size(); ++i) {
        double N_el = 0.0;
        int j = atoms[i].nshl();
        while (N_el < pt[(size_t)atoms[i].Zat].nvb && j > 0) { N_el += atoms[i].Nel[(size_t)j - 1]; --j; }
        n_core[i] = (size_t)j;
        N_e_VB += N_el * atoms[i].Pers;
    }
    Matter.N_VB_el = N_e_VB;
    for (size_t i = 0; i < atoms.size(); ++i) {
        Atom &a = atoms[i];
        const size_t Shl = n_core[i] + (i == 0 ? 1 : 0);
        a.Shell_name.resize(Shl); a.Shl_num.resize(Shl); a.Nel.resize(Shl); a.Ip.resize(Shl); a.Ek.resize(Shl); a.Auger.resize(Shl);
        a.Radiat.resize(Shl); a.PQN.resize(Shl); a.KOCS.resize(Shl); a.KOCS_SHI.resize(Shl);
        a.Ritchi.assign(Shl, CDFosc{});
        if (i == 0) {
            const size_t v = Shl - 1;
            a.Shell_name[v] = "Valence"; a.Shl_num[v] = 63; a.Nel[v] = N_e_VB; a.Ip[v] = Matter.Egap; a.Ek[v] = 0.0;
            a.Auger[v] = 1.0e26; a.Radiat[v] = 1.0e27;
        }
    }
}

// reading_material_parameters, Reading_files_and_parameters.f90:1185-1644: full-CDF files, and (with EADL2023.ALL at hand) files
// that give a chemical formula and / or leave the shells to the atomic database (single-pole CDF, VALENCE / PHONON keywords)
bool read_cdf(const std::string &dir, Case &c, const std::vector<PT> &pt, std::string &err) {
    std::string name = c.numpar.CDF_file.empty() ? (c.Material_name + ".cdf") : c.numpar.CDF_file;
    std::string path = dir + "/INPUT_CDF/" + name;
    Lines f;
    if (!f.load(path)) { err = "File " + path + " is not found!"; return false; }
    c.numpar.CDF_file = "INPUT_CDF/" + name;
    std::vector<std::string> t;
    std::string s;
    auto bad = [&](const std::string &w) { err = "Problem reading " + path + ": " + w; return false; };
    // name line: text before '!' or TAB
    if (!f.read_line(s)) return bad("empty file");
    {
        std::string nm = s;
        size_t p = s.find('!');
        if (p != std::string::npos && p >= 1) nm = s.substr(0, p);
        size_t tb = s.find('\t');
        if (tb != std::string::npos) { if (tb >= 1) nm = s.substr(0, tb); else nm = s.substr(1); }
        size_t a = nm.find_first_not_of(" \t"), b = nm.find_last_not_of(" \t");
        c.Matter.Target_name = (a == std::string::npos) ? "" : nm.substr(a, b - a + 1);
    }
    int N = 0;
    if (read_list(f, 1, t) != 0) return bad("number of elements");
    const bool formula = !parse_int(t[0], N);
    if (formula) {                                                   // a chemical formula instead of an element list (:1291-1308)
        c.Matter.Chem = t[0];
        std::string e;
        if (!decompose_compound(t[0], pt, c.atoms, c.warnings, e)) return bad("chemical-formula line: " + e);
        N = (int)c.atoms.size();
    }
    if (N < 1 || N > TRK3_MAX_ATOMS) return bad("unsupported number of elements");
    if (!formula) { c.atoms.assign(N, Atom{}); c.Matter.Chem.clear(); }
    for (int j = 0; j < N && !formula; ++j) {
        Atom &a = c.atoms[j];
        if (read_list(f, 2, t) != 0 || !parse_int(t[0], a.Zat) || !parse_real(t[1], a.Pers)) return bad("element line");
        if (a.Zat <= 0 || a.Zat >= (int)pt.size() || pt[a.Zat].mass <= 0) return bad("unknown element");
        a.Name = pt[a.Zat].name; a.Full_Name = pt[a.Zat].full; a.Mass = pt[a.Zat].mass;
        char buf[64] = "";
        if (std::fabs(a.Pers - 1.0) > 1.0e-6) std::snprintf(buf, sizeof buf, "%.2f", a.Pers);   // write(temp,'(f10.2)')
        c.Matter.Chem += a.Name + buf;
    }
    // density, speed of sound, Fermi energy [, gap]: formatted read of one line, then list-directed from it
    if (!f.read_line(s)) return bad("density line");
    {
        auto tk = tokens(s);
        double v[4]; int got = 0;
        for (size_t i = 0; i < tk.size() && got < 4; ++i) { if (parse_real(tk[i], v[got])) ++got; else break; }
        if (got >= 4) { c.Matter.Dens = v[0]; c.Matter.Vsound = v[1]; c.Matter.E_F = v[2]; c.Matter.Egap = v[3]; }
        else if (got == 3) { c.Matter.Dens = v[0]; c.Matter.Vsound = v[1]; c.Matter.E_F = v[2]; c.Matter.Egap = 0.0; }
        else return bad("density line needs 'density v_sound E_fermi [E_gap]' (old-format file?)");
        if (c.Matter.Egap < 1.0e-1) c.Matter.Egap = 1.0e-1;
    }
    {
        double sm = 0, sp = 0;
        for (auto &a : c.atoms) { sm += a.Mass * a.Pers; sp += a.Pers; }
        c.Matter.At_Dens = 1.0e-3 * c.Matter.Dens / (g_Mp * sm / sp);
        c.Matter.v_f = std::sqrt(2.0 * c.Matter.E_F / g_me);
    }
    // EADL2023.ALL is optional here (the reference refuses to start without it, Check_EPICS_files :867-892)
    Eadl eadl;
    bool have_eadl = false;
    {
        const std::string p = dir + "/INPUT_EADL/EADL2023.ALL";
        if (file_exists(p)) {
            std::string e;
            if (!eadl.load(p, e)) return bad(e);
            have_eadl = true;
        }
    }
    // number of shells of the first element, or no number: the shells come from the atomic database (single-pole CDF)
    if (!f.read_line(s)) s.clear(), f.pos = f.l.size();
    {
        auto tk = tokens(s); int Shl;
        if (tk.empty() || !parse_int(tk[0], Shl)) {
            // SP_CDF branch (:1349-1488): single-pole CDFs for every shell, optionally a user-given valence-band CDF (VALENCE) and
            // phonon CDF (PHONON)
            c.numpar.kind_of_CDF = 1; c.numpar.kind_of_CDF_ph = 1; c.numpar.VB_CDF_defined = false;
            CDFosc vb;
            if (!f.eof() || !tk.empty()) f.backspace();
            while (f.read_line(s)) {
                size_t a0 = s.find_first_not_of(" \t");
                if (a0 == std::string::npos) continue;
                const std::string key = s.substr(a0, 3);
                if (key == "VAL" || key == "Val" || key == "val") {
                    int ncdf = 0, shl_num = 0; double r3[3];
                    c.numpar.VB_CDF_defined = true;
                    if (read_list(f, 5, t) == 0 && parse_int(t[0], ncdf) && parse_int(t[1], shl_num) && parse_real(t[2], r3[0]) && parse_real(t[3], r3[1]) && parse_real(t[4], r3[2]) && ncdf > 0) {
                        vb.E0.resize((size_t)ncdf); vb.A.resize((size_t)ncdf); vb.Gamma.resize((size_t)ncdf);
                        for (int l = 0; l < ncdf; ++l)
                            if (read_list(f, 3, t) != 0 || !parse_real(t[0], vb.E0[(size_t)l]) || !parse_real(t[1], vb.A[(size_t)l]) || !parse_real(t[2], vb.Gamma[(size_t)l])) {
                                c.warnings.push_back("Could not interprete VB CDF parameters. Using single-pole approximation."); c.numpar.VB_CDF_defined = false; break;
                            }
                    } else { c.warnings.push_back("Could not interprete VB CDF parameters. Using single-pole approximation."); c.numpar.VB_CDF_defined = false; }
                } else if (key == "PHO" || key == "Pho" || key == "pho") {
                    int ncdf = 0;
                    c.numpar.kind_of_CDF_ph = 0;
                    if (read_list(f, 1, t) != 0 || !parse_int(t[0], ncdf) || ncdf < 1) { c.numpar.kind_of_CDF_ph = 1; continue; }
                    c.CDF_Phonon.E0.resize((size_t)ncdf); c.CDF_Phonon.A.resize((size_t)ncdf); c.CDF_Phonon.Gamma.resize((size_t)ncdf);
                    for (int l = 0; l < ncdf; ++l)
                        if (read_list(f, 3, t) != 0 || !parse_real(t[0], c.CDF_Phonon.E0[(size_t)l]) || !parse_real(t[1], c.CDF_Phonon.A[(size_t)l]) || !parse_real(t[2], c.CDF_Phonon.Gamma[(size_t)l])) {
                            c.warnings.push_back("Could not interprete phonon CDF parameters. Using single-pole approximation."); c.numpar.kind_of_CDF_ph = 1; break;
                        }
                }
            }
            if (!have_eadl)
                return bad("files that leave their shells to the atomic database (single-pole CDF; chemical-formula / VALENCE / PHONON keyword form) "
                           "need INPUT_EADL/EADL2023.ALL (check_atomic_parameters, Dealing_with_EADL.f90:374), which is absent");
            for (auto &a : c.atoms) { std::string e; if (!eadl_all_shells(eadl, a, c.numpar.include_photons, e)) return bad(e); }
            make_valence_band(c.atoms, pt, c.Matter);
            for (size_t j = 0; j < c.atoms.size(); ++j) {
                Atom &a = c.atoms[j];
                if (a.nshl() > TRK3_MAX_SHELLS) return bad("too many shells");
                for (int k = 0; k < a.nshl(); ++k) {
                    a.KOCS[(size_t)k] = 1; a.KOCS_SHI[(size_t)k] = 1;
                    if (j == 0 && k == a.nshl() - 1 && c.numpar.VB_CDF_defined) a.Ritchi[(size_t)k] = vb;
                    else { a.Ritchi[(size_t)k].E0.assign(1, 0.0); a.Ritchi[(size_t)k].A.assign(1, 0.0); a.Ritchi[(size_t)k].Gamma.assign(1, 0.0); }
                    if (a.Ip[(size_t)k] < 1.0e-1 && !(j == 0 && k == a.nshl() - 1)) a.Ip[(size_t)k] = a.Ip[(size_t)k];      // (the SP branch keeps EADL's values as they are)
                }
            }
            if (c.numpar.kind_of_CDF_ph == 1) { c.CDF_Phonon.E0.assign(1, 0.0); c.CDF_Phonon.A.assign(1, 0.0); c.CDF_Phonon.Gamma.assign(1, 0.0); }
            return true;
        }
        f.backspace();
    }
    c.numpar.kind_of_CDF = 0;
    for (int j = 0; j < N; ++j) {
        Atom &a = c.atoms[j];
        int Shl;
        if (read_list(f, 1, t) != 0 || !parse_int(t[0], Shl) || Shl < 0) return bad("number of shells");   // 0: an atom without shells of its own (H in H2O.cdf: its electron sits in the valence band of atom 1)
        a.Shell_name.assign(Shl, ""); a.Shl_num.assign(Shl, 0); a.Nel.assign(Shl, 0.0); a.Ip.assign(Shl, -1.0e15);
        a.Ek.assign(Shl, -1.0e-15); a.Auger.assign(Shl, 1.0e31); a.Radiat.assign(Shl, 2.0e31); a.PQN.assign(Shl, 0);
        a.KOCS.assign(Shl, 0); a.KOCS_SHI.assign(Shl, 0); a.Ritchi.assign(Shl, CDFosc{});
        for (int k = 0; k < Shl; ++k) {
            int ncdf, des;
            if (read_list(f, 5, t) != 0 || !parse_int(t[0], ncdf) || !parse_int(t[1], des) || !parse_real(t[2], a.Ip[k]) ||
                !parse_real(t[3], a.Nel[k]) || !parse_real(t[4], a.Auger[k])) return bad("shell line");
            a.Shl_num[k] = std::abs(des);
            define_PQN(a.Shl_num[k], a.Shell_name[k], a.PQN[k]);
            if (a.Shl_num[k] >= 63) c.Matter.N_VB_el = a.Nel[k];
            if (have_eadl) {
                // check_atomic_parameters (Dealing_with_EADL.f90:325-372): what the line left out comes from EADL2023.ALL
                eadl_check_shell(eadl, a, k, c.numpar.include_photons, c.warnings);
                if (a.Nel[k] <= 0 || a.Ip[k] <= -1.0e-14) return bad("shell without Nel/Ip that EADL2023.ALL does not list either");
                if (a.Shl_num[k] < 63 && (a.Auger[k] <= 0.0 || a.Auger[k] > 1.0e30)) return bad("shell without Auger time that EADL2023.ALL does not list either");
                if (a.Radiat[k] == 2.0e31) a.Radiat[k] = 1.0e23;  // element not in the file at all: channel closed
            } else {
                // the same decisions without the EPICS files: only the decay times are decided here, Nel/Ip/Auger must
                // come from the .cdf and the radiative widths from the side-car (apply_radiative_data)
                if (a.Nel[k] <= 0 || a.Ip[k] <= -1.0e-14) return bad("shell without Nel/Ip needs INPUT_EADL/EADL2023.ALL (absent)");
                if (a.Shl_num[k] >= 63 || !c.numpar.include_photons) a.Radiat[k] = 1.0e23;
                else a.Radiat[k] = -1.0;                       // resolved by apply_radiative_data()
                if (a.Shl_num[k] >= 63) a.Auger[k] = 1.0e23;
                else if (a.Auger[k] <= 0.0 || a.Auger[k] > 1.0e30) return bad("shell without Auger time needs INPUT_EADL/EADL2023.ALL (absent)");
            }
            if (a.Ip[k] < 1.0e-1) a.Ip[k] = 1.0e-1;
            if (ncdf > 0) {
                a.KOCS_SHI[k] = 1;
                a.KOCS[k] = (des > 0) ? 1 : 2;
                CDFosc &o = a.Ritchi[k];
                o.E0.resize(ncdf); o.A.resize(ncdf); o.Gamma.resize(ncdf);
                for (int l = 0; l < ncdf; ++l)
                    if (read_list(f, 3, t) != 0 || !parse_real(t[0], o.E0[l]) || !parse_real(t[1], o.A[l]) || !parse_real(t[2], o.Gamma[l]))
                        return bad("CDF oscillator line");
            } else { a.KOCS[k] = 2; a.KOCS_SHI[k] = 2; }
            if (a.KOCS_SHI[k] != 1)
                return bad("a shell without CDF oscillators would need the ion's BEB branch, which the reference itself marks 'NOT WORKING FOR SHI' "
                           "(Monte_Carlo.f90:1774), and EPDL photo-absorption: not supported (BEB shell)");
            if (a.KOCS[k] == 2 && !(a.Ek[k] > 0.0))
                return bad("BEB shell (negative designator): the mean kinetic energy of the shell comes from INPUT_EADL/EADL2023.ALL (I = 914), which is absent");
        }
    }
    // phonon peaks (optional)
    int ncdf = 0;
    int reason = read_list(f, 1, t);
    if (reason != 0) { c.numpar.kind_of_CDF_ph = 1; }
    else if (!parse_int(t[0], ncdf)) { c.numpar.kind_of_CDF_ph = 1; return bad("phonon block"); }
    else {
        c.CDF_Phonon.E0.resize(ncdf); c.CDF_Phonon.A.resize(ncdf); c.CDF_Phonon.Gamma.resize(ncdf);
        for (int l = 0; l < ncdf; ++l)
            if (read_list(f, 3, t) != 0 || !parse_real(t[0], c.CDF_Phonon.E0[l]) || !parse_real(t[1], c.CDF_Phonon.A[l]) || !parse_real(t[2], c.CDF_Phonon.Gamma[l]))
                return bad("phonon oscillator line");
        c.numpar.kind_of_CDF_ph = 0;
    }
    if (c.numpar.kind_of_CDF_ph == 1) { c.CDF_Phonon.E0.assign(1, 0.0); c.CDF_Phonon.A.assign(1, 0.0); c.CDF_Phonon.Gamma.assign(1, 0.0); }
    return true;
}

// Radiative decay times (Dealing_with_EADL.f90:350-360: t[fs] = 1e15*hbar/(e*Gamma_R[eV]); Gamma_R<1e-6 => 1.1e35).
// EADL2023.ALL is not redistributable with the reference tree, so the widths are taken from an optional side-car
//   INPUT_EADL/radiative_widths.dat :  "Z  designator  Gamma_R[eV]"  per line ('!' comments)
// If a shell has no entry, the channel is closed (1e23 fs, the reference's value for "not included") and a warning is kept.
void apply_radiative_data(const std::string &dir, Case &c) {
    if (!c.numpar.include_photons) return;
    struct W { int z, d; double g; };
    std::vector<W> w;
    Lines f;
    if (f.load(dir + "/INPUT_EADL/radiative_widths.dat")) {
        for (auto &s : f.l) {
            auto t = tokens(s); W x;
            if (t.size() >= 3 && parse_int(t[0], x.z) && parse_int(t[1], x.d) && parse_real(t[2], x.g)) w.push_back(x);
        }
    }
    for (auto &a : c.atoms)
        for (int k = 0; k < a.nshl(); ++k) {
            if (a.Radiat[k] >= 0.0) continue;
            bool found = false;
            for (auto &x : w) if (x.z == a.Zat && x.d == a.Shl_num[k]) {
                a.Radiat[k] = (x.g < 1.0e-6) ? 1.1e35 : 1.0e15 * g_h / (g_e * x.g);
                found = true; break;
            }
            if (!found) {
                a.Radiat[k] = 1.0e23;
                c.warnings.push_back("no radiative width for Z=" + std::to_string(a.Zat) + " shell designator " + std::to_string(a.Shl_num[k]) +
                                     " (EADL2023.ALL / radiative_widths.dat absent): radiative channel closed for this shell");
            }
        }
}

// Linear_approx_2d(Array, In_val, Value1, El1, El2), Reading_files_and_parameters.f90:3272
double linear_approx_2d(const std::vector<double> &X, const std::vector<double> &Y, double v, double El1, double El2) {
    int N = (int)X.size();
    int num = find_monoton_2d(X.data(), 1, N, v);
    if (num == 1) return El2 + (Y[0] - El2) / (X[0] - El1) * (v - El1);
    if (Y[num - 2] > 1e20) return Y[num - 2];
    return Y[num - 2] + (Y[num - 1] - Y[num - 2]) / (X[num - 1] - X[num - 2]) * (v - X[num - 2]);
}

// reading_material_DOS, Reading_files_and_parameters.f90:2317-2454
bool read_dos(const std::string &dir, Case &c, std::string &err) {
    std::string name = c.numpar.DOS_file.empty() ? (c.Material_name + ".dos") : c.numpar.DOS_file;
    std::string path = dir + "/INPUT_DOS/" + name;
    if (!file_exists(path)) {
        std::string fr = dir + "/INPUT_DOS/Free_electron_DOS.dos";
        if (!file_exists(fr)) { err = "Files " + path + " and " + fr + " are not found!"; return false; }
        path = fr; name = "Free_electron_DOS.dos";
    }
    c.numpar.DOS_file = "INPUT_DOS/" + name;
    Lines f;
    if (!f.load(path)) { err = "cannot open " + path; return false; }
    std::vector<double> X, Y;
    // Count_lines_in_file counts records (list-directed empty reads); every record must then hold two reals
    for (auto &s : f.l) {
        auto t = tokens(s);
        double a, b;
        if (t.size() < 2 || !parse_real(t[0], a) || !parse_real(t[1], b)) {
            if (t.empty()) continue;      // trailing blank record: a list-directed read would hit EOF instead
            err = "Problem reading " + path; return false;
        }
        X.push_back(a); Y.push_back(b);
    }
    int N = (int)X.size();
    if (N < 2) { err = "DOS file too short: " + path; return false; }
    double top = X[N - 1];
    for (auto &x : X) x = x - top;
    const double dE = 0.1;
    int M = (int)std::floor((X[N - 1] - X[0]) / dE);
    DOS &d = c.dos;
    d.E.assign(M, 0); d.dos.assign(M, 0); d.int_DOS.assign(M, 0); d.k.assign(M, 0); d.Eff_m.assign(M, 0);
    d.DOS_inv.assign(M, 0); d.int_DOS_inv.assign(M, 0); d.k_inv.assign(M, 0); d.Eff_m_inv.assign(M, 0);
    double E = X[0];
    for (int i = 0; i < M; ++i) {
        E = E + dE;
        d.E[i] = E;
        d.dos[i] = linear_approx_2d(X, Y, E, X[0] - dE, 0.0);
    }
    double Elast = d.E[M - 1];
    for (auto &x : d.E) x = std::fabs(x - Elast);
    std::reverse(d.E.begin(), d.E.end());
    std::reverse(d.dos.begin(), d.dos.end());
    d.DOS_inv = d.dos; std::reverse(d.DOS_inv.begin(), d.DOS_inv.end());
    double s = 0, si = 0;
    for (int i = 0; i < M; ++i) { s += d.dos[i]; d.int_DOS[i] = s; si += d.DOS_inv[i]; d.int_DOS_inv[i] = si; }
    double SUM = d.int_DOS[M - 1];
    for (int i = 0; i < M; ++i) { d.dos[i] = d.dos[i] / SUM * c.Matter.N_VB_el; d.int_DOS[i] = d.int_DOS[i] / SUM * c.Matter.N_VB_el; }
    double spers = 0; for (auto &a : c.atoms) spers += a.Pers;
    s = 0;
    for (int i = 0; i < M; ++i) {
        s += d.dos[i];
        d.k[i] = std::pow(3.0 * 2.0 * g_Pi * g_Pi / 2.0 * s * c.Matter.At_Dens / spers * 1e6, 1.0 / 3.0);
        if (d.E[i] < 1.0e-10) d.Eff_m[i] = 1.0;
        else d.Eff_m[i] = g_h * g_h * d.k[i] * d.k[i] / (2.0 * d.E[i] * g_e) / g_me;
    }
    SUM = d.int_DOS_inv[M - 1];
    for (int i = 0; i < M; ++i) { d.DOS_inv[i] = d.DOS_inv[i] / SUM * c.Matter.N_VB_el; d.int_DOS_inv[i] = d.int_DOS_inv[i] / SUM * c.Matter.N_VB_el; }
    for (int i = 0; i < M; ++i) {
        d.k_inv[i] = std::pow(3.0 * 2.0 * g_Pi * g_Pi / 2.0 * d.int_DOS_inv[i] * c.Matter.At_Dens / spers * 1e6, 1.0 / 3.0);
        if (d.E[i] < 1.0e-10) d.Eff_m_inv[i] = 1.0;
        else d.Eff_m_inv[i] = g_h * g_h * d.k_inv[i] * d.k_inv[i] / (2.0 * d.E[i] * g_e) / g_me;
    }
    return true;
}

}  // namespace

void set_default_numpar(NumPar &np) { np = NumPar{}; }

bool read_case(const std::string &dir, Case &c, std::string &err) {
    c = Case{};
    c.dir = dir;
    set_default_numpar(c.numpar);
    std::vector<PT> pt;
    if (!load_periodic_table(dir, pt, err)) return false;
    if (!read_input_parameters(dir, c, pt, err)) return false;
    if (!read_cdf(dir, c, pt, err)) return false;
    apply_radiative_data(dir, c);
    if (c.SHI.Kind_ion != 0 && c.SHI.Kind_ion != 1) c.SHI.Kind_ion = 0;       // anything but 1 is the point charge (select case default)
    if (c.numpar.kind_of_DR == 4) {
        // Delta-CDF: the weights follow the oscillators as they are read (Reading_files_and_parameters.f90:1565-1575, 1608-1610)
        if (c.numpar.kind_of_CDF == 1) { err = "Delta-CDF (kind_of_DR=4) with a .cdf that leaves its shells to the atomic database (single-pole CDFs) is not supported: "
                                              "the reference allocates the delta-function weights of such shells and never assigns them (Reading_files_and_parameters.f90:1482-1484)"; return false; }
        for (auto &a : c.atoms) for (int k = 0; k < a.nshl(); ++k) {
            CDFosc &o = a.Ritchi[(size_t)k];
            o.alpha.resize(o.E0.size());
            for (size_t l = 0; l < o.E0.size(); ++l) o.alpha[l] = define_alpha(o.A[l], o.Gamma[l], o.E0[l], a.Ip[(size_t)k]);
        }
        CDFosc &ph = c.CDF_Phonon;
        ph.alpha.resize(ph.E0.size());
        for (size_t l = 0; l < ph.E0.size(); ++l) ph.alpha[l] = define_alpha(ph.A[l], ph.Gamma[l], ph.E0[l], 0.0);
    }
    if (c.numpar.kind_of_EMFP == 2) {
        // Reading_files_and_parameters.f90:578-599: electrons, then holes; a missing file switches the run to Mott cross sections
        for (int hole = 0; hole < 2 && c.numpar.kind_of_EMFP == 2; ++hole) {
            const std::string rel = dsf_file_name(c, hole != 0);
            bool found = false;
            if (!read_dsf(dir + "/" + rel, hole ? c.DSF_DEMFP_H : c.DSF_DEMFP, found, err)) return false;
            if (!found) {
                c.warnings.push_back("File " + rel + " is not found. The calculations proceed with Mott atomic cross-sections.");
                c.numpar.kind_of_EMFP = 0; c.DSF_DEMFP.clear(); c.DSF_DEMFP_H.clear();
            }
        }
    }
    if (c.numpar.CDF_elast_Zeff == 2) {
        // read_form_factors, Reading_files_and_parameters.f90:605-615, 867-907: one header line, then row Z = a1..a5 of element Z
        const std::string p = dir + "/INPUT_EADL/Atomic_form_factors.dat";
        std::ifstream f(p);
        if (!f) { err = "CDF_elast_Zeff=2 needs " + p; return false; }
        std::string line;
        std::getline(f, line);
        while (std::getline(f, line)) {
            std::istringstream is(line);
            std::array<double, 5> a{};
            if (!(is >> a[0] >> a[1] >> a[2] >> a[3] >> a[4])) { if (line.find_first_not_of(" \t\r") == std::string::npos) continue; err = "could not read a line of " + p; return false; }
            c.form_factor.push_back(a);
        }
        for (auto &a : c.atoms) if (a.Zat < 1 || a.Zat > (int)c.form_factor.size()) { err = "no form factor for Z=" + std::to_string(a.Zat) + " in " + p; return false; }
    }
    if (c.numpar.CDF_elast_Zeff == 3) {
        // get_screening via construct_CDF(..., 1, size(Target_atoms(i)%Ip), ...) (Cross_sections.f90:3253) addresses shell number
        // "shells of atom i" of the FIRST atom: out of bounds in the reference when another atom has more shells than the first
        for (auto &a : c.atoms) if (a.nshl() > c.atoms[0].nshl()) { err = "CDF_elast_Zeff=3: an atom has more shells than the first atom (the reference reads out of bounds)"; return false; }
    }
    if (c.numpar.CDF_elast_Zeff > 3 || c.numpar.CDF_elast_Zeff < 0) c.numpar.CDF_elast_Zeff = 0;   // select case default: Barkas-like charge
    if (c.numpar.CS_method != 1) { err = "only 'grid 1' (tabulated differential cross sections, the reference default) is supported"; return false; }
    if (c.SHI.Zat > 0) {
        // Reading_files_and_parameters.f90:509-535
        const Atom &a1 = c.atoms[0];
        double M = c.SHI.Mass * g_Mp;
        double Emin = (M + g_me) * (M + g_me) / (M * g_me) * a1.Ip.back() / 4.0;
        if (c.SHI.E <= Emin) { err = "The SHI energy is smaller than the minimum allowed energy"; return false; }
        if (c.SHI.E >= 175.0e6 / 2.0 * c.SHI.Mass) { err = "The SHI energy is higher than the maximum allowed energy"; return false; }
    }
    if (!read_dos(dir, c, err)) return false;
    return true;
}

}  // namespace trk3
