// boost::to_upper_copy stand-in (viterbi_alignment.cpp:103). TEST INFRASTRUCTURE.
#ifndef PAGAN2_B200_SHIM_ALGORITHM_STRING_HPP
#define PAGAN2_B200_SHIM_ALGORITHM_STRING_HPP
#include <string>
#include <cctype>
namespace boost {
inline std::string to_upper_copy(const std::string &s) {
    std::string r(s);
    for (size_t i = 0; i < r.size(); i++) r[i] = (char)std::toupper((unsigned char)r[i]);
    return r;
}
}
#endif
