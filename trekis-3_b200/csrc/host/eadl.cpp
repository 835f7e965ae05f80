// eadl.cpp -- reader of the EPICS atomic-relaxation library EADL2023.ALL (ENDL format) for the atomic parameters a .cdf
// file leaves out (SURVEY.md Appendix C, row N4 of 8(f)): check_atomic_parameters, READ_EADL_TYPE_FILE_int/_real,
// select_imin_imax, next_designator (Dealing_with_EADL.f90:312-742, 913-938).
//
// ENDL layout as the reference reads it: a block is two header lines, data lines, and a terminator line with '1' in
// column 72 (How_many_lines :927-938, format 71X,I1).  Header 1 starts with Z in columns 1-3 (:933 format I3,I3,I2,I2,E11.4,I6),
// header 2 with C (columns 1-2), I (3-5), S (6-8) (format I2,I3,I3,E11.4).  Data lines of the blocks used here are read
// list-directed as (shell designator, value); energies and widths are in MeV.
//   I = 912 electrons per subshell, 913 binding energy, 914 kinetic energy, 921 radiative width, 922 non-radiative width.
// The file itself (IAEA EPICS2023) is not part of the reference tree; the reference refuses to start without it.  Here it is
// optional: when <dir>/INPUT_EADL/EADL2023.ALL exists it is used exactly where the reference uses it, otherwise the .cdf
// must be complete and the radiative widths come from the side-car radiative_widths.dat (input.cpp).
#include <cmath>
#include <fstream>
#include "fortran_fmt.hpp"
#include "trk3_host.hpp"

namespace trk3 {

bool Eadl::load(const std::string &path, std::string &err) {
    std::ifstream f(path);
    if (!f) { err = "File " + path + " is not found!"; return false; }
    std::vector<std::string> lines;
    std::string s;
    while (std::getline(f, s)) { if (!s.empty() && s.back() == '\r') s.pop_back(); lines.push_back(s); }
    auto field = [](const std::string &l, size_t a, size_t n) { int v = 0; if (l.size() > a) v = atoi(l.substr(a, n).c_str()); return v; };
    blocks.clear();
    size_t i = 0;
    while (i + 1 < lines.size()) {
        Block b;
        b.Z = field(lines[i], 0, 3);
        b.C = field(lines[i + 1], 0, 2); b.I = field(lines[i + 1], 2, 3); b.S = field(lines[i + 1], 5, 3);
        i += 2;
        bool closed = false;
        for (; i < lines.size(); ++i) {
            const std::string &l = lines[i];
            if (l.size() >= 72 && l[71] == '1') { closed = true; ++i; break; }
            std::vector<double> v;
            if (parse_fortran_line(l, v) && v.size() >= 2) b.rows.push_back({v[0], v[1]});
            else b.rows.push_back({std::nan(""), std::nan("")});       // a line the list-directed read would choke on
        }
        if (!closed && b.rows.empty()) break;
        if (b.Z > 0) blocks.push_back(std::move(b));
    }
    if (blocks.empty()) { err = path + " holds no ENDL blocks"; return false; }
    return true;
}

const Eadl::Block *Eadl::find(int Z, int I) const {
    // first block of the element, then forward until the reaction matches or the next element starts (:578-596)
    for (const Block &b : blocks) {
        if (b.Z > Z) return nullptr;
        if (b.Z == Z && b.I == I) return &b;
    }
    return nullptr;
}
bool Eadl::has_element(int Z) const { for (const Block &b : blocks) if (b.Z == Z) return true; return false; }

void eadl_select_imin_imax(int des, int &imin, int &imax) {      // :687-722
    switch (des) {
    case 2: imin = 3; imax = 6; break;      case 4: imin = 5; imax = 6; break;
    case 7: imin = 8; imax = 14; break;     case 9: imin = 10; imax = 11; break;
    case 12: imin = 13; imax = 14; break;   case 15: imin = 16; imax = 25; break;
    case 26: imin = 27; imax = 39; break;   case 40: imin = 41; imax = 56; break;
    case 57: imin = 58; imax = 61; break;   default: imin = 0; imax = 0;
    }
}
int eadl_next_designator(int last) {                               // :725-742
    if (last <= 1) return 2;
    if (last <= 6) return 7;
    if (last <= 14) return 15;
    if (last <= 25) return 26;
    if (last <= 39) return 40;
    return 57;
}

// READ_EADL_TYPE_FILE_real, single-shell branch (:599-660).  The value of the designator if the block lists it; otherwise the
// reference re-reads the block from its first line, (imax-imin+1) lines at most, and AVERAGES the values of the lines whose
// designator is >= imin until it meets imax -- for a whole shell with more sub-shells than that window covers (M, N, O...)
// only the first ones count.  Kept as it is: it decides the decay times the Monte-Carlo runs with.
bool Eadl::real_value(int Z, int I, int des, double &out) const {
    if (!has_element(Z)) return false;                             // the reference leaves the array untouched
    const Block *b = find(Z, I);
    if (!b) { out = 1.0e-30; return true; }                        // :604-607
    if (des >= 63) { out = 1.0e22; return true; }                  // :662
    for (const Row &r : b->rows) if (r.des == (double)des) { out = r.val * 1.0e6; if (out != out) out = 0.0; return true; }
    int imin, imax;
    eadl_select_imin_imax(des, imin, imax);
    double sum = 0.0; int icont = 0;
    size_t line = 0;
    for (int run = imin; run <= imax && line < b->rows.size(); ++run, ++line) {
        const Row &r = b->rows[line];
        if (r.des >= imin) { if (r.val == r.val) sum += r.val; ++icont; }
        if (r.des == (double)imax) break;
    }
    out = icont > 0 ? (sum / (double)icont) * 1.0e6 : 0.0;
    if (out != out) out = 0.0;
    return true;
}

// READ_EADL_TYPE_FILE_int, single-shell branch (:459-505): electrons of the designator, or the SUM over the same window
bool Eadl::electrons(int Z, int des, double &out) const {
    if (des >= 63) return false;
    const Block *b = find(Z, 912);
    if (!b) return false;
    for (const Row &r : b->rows) if (r.des == (double)des) { out = r.val; return true; }
    int imin, imax;
    eadl_select_imin_imax(des, imin, imax);
    double sum = 0.0;
    size_t line = 0;
    for (int run = imin; run <= imax && line < b->rows.size(); ++run, ++line) {
        const Row &r = b->rows[line];
        if (r.des >= imin) sum += r.val;
        if (r.des == (double)imax) break;
    }
    out = sum;
    return true;
}

// check_atomic_parameters, single-shell branch (:325-372), for shell k of atom a (whose Shl_num, Nel, Ip, Ek, Auger, Radiat hold
// what the .cdf gave)
void eadl_check_shell(const Eadl &db, Atom &a, int k, bool include_photons, std::vector<std::string> &warnings) {
    const int Z = a.Zat, des = a.Shl_num[(size_t)k];
    if (!db.has_element(Z)) { warnings.push_back("element Z=" + std::to_string(Z) + " is not in the EADL database"); return; }
    double v;
    if (a.Nel[k] <= 0 && db.electrons(Z, des, v)) a.Nel[k] = v;                                   // :327-329
    if (a.Ip[k] <= -1.0e-14 && db.real_value(Z, 913, des, v)) a.Ip[k] = v;                        // :330-332
    if (a.Ek[k] <= 0.0) {                                                                           // :333-343
        const int d = des >= 62 ? eadl_next_designator(k > 0 ? a.Shl_num[(size_t)k - 1] : 0) : des;
        if (db.real_value(Z, 914, d, v)) a.Ek[k] = v;
    }
    if (des >= 63 || !include_photons) a.Radiat[k] = 1.0e23;                                        // :344-346
    else if (db.real_value(Z, 921, des, v)) a.Radiat[k] = v < 1.0e-6 ? 1.1e35 : 1.0e15 * g_h / (g_e * v);   // :347-354
    if (des >= 63) a.Auger[k] = 1.0e23;                                                             // :358-360
    else if ((a.Auger[k] <= 0.0 || a.Auger[k] > 1.0e30) && db.real_value(Z, 922, des, v)) a.Auger[k] = 1.0e15 * g_h / (g_e * v);   // :361-363
}

}  // namespace trk3
