// Empty stand-in: the reference only names boost::multi_array in comments (viterbi_alignment.h:49-59).
#ifndef PAGAN2_B200_SHIM_MULTI_ARRAY_HPP
#define PAGAN2_B200_SHIM_MULTI_ARRAY_HPP
#endif
