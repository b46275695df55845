// The link flags of a handle as a device-side constant.  One image per flag combination
// (-DB200_LINK_FLAGS=k) is embedded in the library; b200ode_link_rhs() adds the matching one
// to the nvJitLink link, where LTO inlines the function, so `if (b200ode_link_flags() & ...)`
// in a solver kernel is resolved at link time and the branch not taken leaves no code.
#ifndef B200_LINK_FLAGS
#error "compile with -DB200_LINK_FLAGS=<flags>"
#endif
extern "C" __device__ unsigned b200ode_link_flags() { return B200_LINK_FLAGS; }
