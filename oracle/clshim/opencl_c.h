/* TEST INFRASTRUCTURE -- not part of the product path.
 *
 * A minimal "OpenCL C on gcc" translation layer: with this header force-included
 * (`gcc -x c -include opencl_c.h`), an OpenCL C kernel file that only uses address-space
 * qualifiers, work-item functions and barrier() compiles as plain C11.  It exists so that
 * the reference's own kernel sources (LB_D2Q9/D2Q9.cl, D2Q9i.cl) can be compiled where they
 * lie under /root/reference and executed on the CPU, work-item by work-item, as the ground
 * truth the restated oracle is pinned against (oracle/build_ref.py, oracle/shims/pyopencl).
 *
 * Semantics kept: `float` arithmetic stays float, double literals promote exactly as in
 * OpenCL C (same usual arithmetic conversions as C), no contraction (-ffp-contract=off).
 */
#ifndef CLSHIM_OPENCL_C_H
#define CLSHIM_OPENCL_C_H
#include <stddef.h>
#include <stdint.h>
#include <tgmath.h>          /* OpenCL's math built-ins are type-generic */

#define __kernel
#define __global
#define __local
#define __constant
#define __private
#define __read_only
#define __write_only
#define __read_write

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef unsigned long ulong;

#define cl_khr_fp64 1          /* what an fp64-capable device defines; multi.cl tests for it */
#define CLK_LOCAL_MEM_FENCE 1
#define CLK_GLOBAL_MEM_FENCE 2

typedef struct clshim_item {
    size_t gid[3], lid[3], grp[3];      /* global id, local id, group id */
    size_t gsz[3], lsz[3], ngrp[3];     /* global size, local size, number of groups */
    unsigned dim;
} clshim_item;

extern __thread clshim_item *clshim_cur;   /* the work-item being executed by this thread */
void clshim_barrier(void);                 /* yields to the other work-items of the group */

static inline size_t get_global_id(uint d) { return d < 3 ? clshim_cur->gid[d] : 0; }
static inline size_t get_local_id(uint d) { return d < 3 ? clshim_cur->lid[d] : 0; }
static inline size_t get_group_id(uint d) { return d < 3 ? clshim_cur->grp[d] : 0; }
static inline size_t get_global_size(uint d) { return d < 3 ? clshim_cur->gsz[d] : 1; }
static inline size_t get_local_size(uint d) { return d < 3 ? clshim_cur->lsz[d] : 1; }
static inline size_t get_num_groups(uint d) { return d < 3 ? clshim_cur->ngrp[d] : 1; }
static inline size_t get_global_offset(uint d) { (void)d; return 0; }
static inline uint get_work_dim(void) { return clshim_cur->dim; }

#define barrier(flags) clshim_barrier()
#define mem_fence(flags) ((void)0)
#define read_mem_fence(flags) ((void)0)
#define write_mem_fence(flags) ((void)0)

/* entry point of the NDRange executor (ndrange.c); thunk(args) runs ONE work-item */
int clshim_run(unsigned dim, const size_t *gsize, const size_t *lsize,
               void (*thunk)(void *), void *args, int uses_barrier);
#endif
