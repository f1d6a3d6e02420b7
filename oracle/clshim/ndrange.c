/* TEST INFRASTRUCTURE -- NDRange executor for opencl_c.h.
 *
 * Work-groups run one after another; inside a group the work-items run in local-id order.
 * A kernel that calls barrier() gets one fiber per work-item: barrier() switches back to the
 * scheduler, which resumes the next item, so every item of the group reaches barrier k before
 * any item continues past it -- the OpenCL execution model, serialised.  Fibers are a dozen
 * lines of x86-64 (callee-saved registers + stack pointer); other hosts, or
 * -DCLSHIM_USE_UCONTEXT, fall back to <ucontext.h>.
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <string.h>
#include "opencl_c.h"

#if !defined(__x86_64__) && !defined(CLSHIM_USE_UCONTEXT)
#define CLSHIM_USE_UCONTEXT 1
#endif
#ifdef CLSHIM_USE_UCONTEXT
#include <ucontext.h>
#endif

__thread clshim_item *clshim_cur = 0;

#define FIBER_STACK (64 * 1024)

typedef struct fiber {
#ifdef CLSHIM_USE_UCONTEXT
    ucontext_t ctx;
#else
    void *sp;
#endif
    clshim_item item;
    int done;
} fiber;

static __thread fiber *cur_fiber = 0;
static __thread void (*cur_thunk)(void *) = 0;
static __thread void *cur_args = 0;

#ifdef CLSHIM_USE_UCONTEXT
static __thread ucontext_t sched_ctx;
static void to_scheduler(fiber *f) { swapcontext(&f->ctx, &sched_ctx); }
static void to_fiber(fiber *f) { swapcontext(&sched_ctx, &f->ctx); }
static void fiber_entry(void)
{
    fiber *self = cur_fiber;
    cur_thunk(cur_args);
    self->done = 1;                      /* uc_link returns to the scheduler */
}
static void fiber_init(fiber *f, char *stack)
{
    getcontext(&f->ctx);
    f->ctx.uc_stack.ss_sp = stack;
    f->ctx.uc_stack.ss_size = FIBER_STACK;
    f->ctx.uc_link = &sched_ctx;
    makecontext(&f->ctx, fiber_entry, 0);
}
#else
static __thread void *sched_sp;
/* save the callee-saved registers on the current stack, publish its pointer, adopt new_sp */
__attribute__((naked, noinline)) static void ctx_switch(void **save_sp, void *new_sp)
{
    __asm__ volatile(
        "pushq %rbp\n\tpushq %rbx\n\tpushq %r12\n\tpushq %r13\n\tpushq %r14\n\tpushq %r15\n\t"
        "movq %rsp, (%rdi)\n\t"
        "movq %rsi, %rsp\n\t"
        "popq %r15\n\tpopq %r14\n\tpopq %r13\n\tpopq %r12\n\tpopq %rbx\n\tpopq %rbp\n\t"
        "ret\n\t");
}
static void to_scheduler(fiber *f) { ctx_switch(&f->sp, sched_sp); }
static void to_fiber(fiber *f) { ctx_switch(&sched_sp, f->sp); }
static void fiber_entry(void)
{
    fiber *self = cur_fiber;
    cur_thunk(cur_args);
    self->done = 1;
    to_scheduler(self);                  /* never resumed */
    __builtin_trap();
}
static void fiber_init(fiber *f, char *stack)
{
    uintptr_t top = ((uintptr_t)stack + FIBER_STACK) & ~(uintptr_t)15;
    void **sp = (void **)top;
    *--sp = 0;                           /* keeps rsp = 8 (mod 16) at fiber_entry, as after a call */
    *--sp = (void *)fiber_entry;         /* `ret` target of the first switch */
    for (int i = 0; i < 6; ++i) *--sp = 0;
    f->sp = sp;
}
#endif

void clshim_barrier(void)
{
    if (cur_fiber)                       /* barrier-free launches run items to completion */
        to_scheduler(cur_fiber);
}

static void fill_item(clshim_item *it, unsigned dim, const size_t *g, const size_t *l,
                      const size_t *grp, const size_t *lid)
{
    it->dim = dim;
    for (int d = 0; d < 3; ++d) {
        it->gsz[d] = g[d]; it->lsz[d] = l[d]; it->ngrp[d] = g[d] / l[d];
        it->grp[d] = grp[d]; it->lid[d] = lid[d];
        it->gid[d] = grp[d] * l[d] + lid[d];
    }
}

int clshim_run(unsigned dim, const size_t *gsize, const size_t *lsize,
               void (*thunk)(void *), void *args, int uses_barrier)
{
    size_t g[3] = {1, 1, 1}, l[3] = {1, 1, 1};
    if (dim < 1 || dim > 3) return -1;
    for (unsigned d = 0; d < dim; ++d) {
        g[d] = gsize[d];
        l[d] = lsize ? lsize[d] : 1;
        if (l[d] == 0 || g[d] % l[d]) return -2;         /* CL_INVALID_WORK_GROUP_SIZE */
    }
    const size_t per_group = l[0] * l[1] * l[2];
    size_t grp[3], lid[3];

    if (!uses_barrier) {
        clshim_item it;
        clshim_item *saved = clshim_cur;
        clshim_cur = &it;
        for (grp[2] = 0; grp[2] < g[2] / l[2]; ++grp[2])
        for (grp[1] = 0; grp[1] < g[1] / l[1]; ++grp[1])
        for (grp[0] = 0; grp[0] < g[0] / l[0]; ++grp[0])
            for (lid[2] = 0; lid[2] < l[2]; ++lid[2])
            for (lid[1] = 0; lid[1] < l[1]; ++lid[1])
            for (lid[0] = 0; lid[0] < l[0]; ++lid[0]) {
                fill_item(&it, dim, g, l, grp, lid);
                thunk(args);
            }
        clshim_cur = saved;
        return 0;
    }

    fiber *fibers = (fiber *)calloc(per_group, sizeof(fiber));
    char *stacks = (char *)malloc(per_group * (size_t)FIBER_STACK + 64);
    if (!fibers || !stacks) { free(fibers); free(stacks); return -3; }
    cur_thunk = thunk;
    cur_args = args;
    for (grp[2] = 0; grp[2] < g[2] / l[2]; ++grp[2])
    for (grp[1] = 0; grp[1] < g[1] / l[1]; ++grp[1])
    for (grp[0] = 0; grp[0] < g[0] / l[0]; ++grp[0]) {
        size_t n = 0;
        for (lid[2] = 0; lid[2] < l[2]; ++lid[2])
        for (lid[1] = 0; lid[1] < l[1]; ++lid[1])
        for (lid[0] = 0; lid[0] < l[0]; ++lid[0], ++n) {
            fiber *f = &fibers[n];
            f->done = 0;
            fill_item(&f->item, dim, g, l, grp, lid);
            fiber_init(f, stacks + n * (size_t)FIBER_STACK);
        }
        size_t remaining = per_group;
        while (remaining) {              /* one pass = "run every live item to its next barrier" */
            size_t finished = 0;
            for (size_t i = 0; i < per_group; ++i) {
                fiber *f = &fibers[i];
                if (f->done) continue;
                cur_fiber = f;
                clshim_cur = &f->item;
                to_fiber(f);
                if (f->done) ++finished;
            }
            /* divergent barriers (some items finished, others waiting) are undefined in OpenCL;
               the stragglers are simply run to completion */
            remaining -= finished;
        }
    }
    cur_fiber = 0;
    clshim_cur = 0;
    free(fibers);
    free(stacks);
    return 0;
}
