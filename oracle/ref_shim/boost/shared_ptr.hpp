// Empty stand-in: tunnel_matrix.h:26 includes it but uses nothing from it.
#ifndef PAGAN2_B200_SHIM_SHARED_PTR_HPP
#define PAGAN2_B200_SHIM_SHARED_PTR_HPP
#endif
