// output.cpp -- writer of the reference's output directory (Save_output, Sorting_output_data.f90:340-1140).
#include "trk3_host.hpp"
namespace trk3 {
bool save_output(const Case &, const trk3_tally_layout &, const double *, int, const std::string &, std::string &, std::string &err) {
    err = "save_output: not implemented yet";
    return false;
}
}  // namespace trk3
