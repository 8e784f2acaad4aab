/* Stand-in for <sys/sysctl.h>: the reference's optimizations/parallel_fft.c:5 includes it, but
 * glibc >= 2.32 no longer ships it. Only the reference's demo main() (parallel_fft.c:374) calls
 * sysctlbyname; the wrapper never runs that main. Test infrastructure only. */
#ifndef ORACLE_SHIM_SYSCTL_H
#define ORACLE_SHIM_SYSCTL_H
#include <stddef.h>
static inline int sysctlbyname(const char* name, void* oldp, size_t* oldlenp, void* newp, size_t newlen) {
    (void)name; (void)oldp; (void)oldlenp; (void)newp; (void)newlen;
    return -1;
}
#endif
