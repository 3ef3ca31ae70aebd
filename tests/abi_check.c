/* A C consumer of include/satrans_b200.h (tests/test_host_logic.py compiles and runs it): the header is plain C99, and the
 * struct layouts a C binding sees are the ones libsatrans_b200.so was compiled with. */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "satrans_b200.h"

typedef int (*ver_fn)(void);
typedef int (*sizes_fn)(int64_t*);

int main(int argc, char** argv) {
    void* h;
    ver_fn ver;
    sizes_fn sizes;
    int64_t s[8];
    if (argc < 2) return 4;
    h = dlopen(argv[1], RTLD_NOW);
    if (!h) { printf("dlopen failed: %s\n", dlerror()); return 2; }
    *(void**)(&ver) = dlsym(h, "sat_abi_version");
    *(void**)(&sizes) = dlsym(h, "sat_struct_sizes");
    if (!ver || !sizes) return 3;
    memset(s, 0, sizeof(s));
    if (sizes(s) != SAT_OK) return 5;
    printf("abi %d header %d SatModel %lld/%lu SatWork %lld/%lu SatAdam %lld/%lu\n", ver(), SAT_ABI_VERSION, (long long)s[0],
           (unsigned long)sizeof(SatModel), (long long)s[1], (unsigned long)sizeof(SatWork), (long long)s[2],
           (unsigned long)sizeof(SatAdam));
    return (ver() == SAT_ABI_VERSION && s[0] == (int64_t)sizeof(SatModel) && s[1] == (int64_t)sizeof(SatWork) &&
            s[2] == (int64_t)sizeof(SatAdam)) ? 0 : 1;
}
