// fiber.h — the execution contexts of the GICP host loop (gicp_host.inl): the coordinator's and one per scan.
// Host-only, no CUDA: tests/cpp/fiber_test.cpp exercises it on the CPU.
//
// swapcontext() saves and restores the signal mask: two system calls per switch, ~0.3 us each way, which was two thirds
// of the host time of a GICP round.  On x86-64 the switch is six callee-saved registers and the stack pointer (System V
// ABI: everything else is caller-saved across the call to b2_fiber_switch; MXCSR / x87 control words are never changed
// by this code); elsewhere, or with -DB2_FIBER_UCONTEXT, the ucontext path stays.  Include from ONE translation unit
// per binary: the assembly below defines the symbol.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <ucontext.h>

#if defined(__x86_64__) && !defined(B2_FIBER_UCONTEXT) && !defined(__CUDA_ARCH__)
#define B2_FIBER_ASM 1
extern "C" void b2_fiber_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.p2align 4
.globl b2_fiber_switch
.hidden b2_fiber_switch
.type b2_fiber_switch,@function
b2_fiber_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size b2_fiber_switch,.-b2_fiber_switch
)");
#endif

#ifdef B2_FIBER_ASM
struct FiberCtx {
  void* sp = nullptr;
};
inline void fiber_switch(FiberCtx& from, FiberCtx& to) { b2_fiber_switch(&from.sp, to.sp); }
// first switch into `c` "returns" into entry() on the given stack (entry never returns: it switches away for good)
inline void fiber_make(FiberCtx& c, FiberCtx&, char* stack, size_t size, void (*entry)()) {
  void** sp = reinterpret_cast<void**>((reinterpret_cast<uintptr_t>(stack) + size) & ~(uintptr_t)15);
  *--sp = nullptr;                              // where entry's return address would be
  *--sp = reinterpret_cast<void*>(entry);       // popped by the switch's `ret`; entry then sees rsp = 8 mod 16
  for (int i = 0; i < 6; ++i) *--sp = nullptr;  // rbp rbx r12 r13 r14 r15
  c.sp = sp;
}
#else
struct FiberCtx {
  ucontext_t uc;
};
inline void fiber_switch(FiberCtx& from, FiberCtx& to) { swapcontext(&from.uc, &to.uc); }
inline void fiber_make(FiberCtx& c, FiberCtx& back, char* stack, size_t size, void (*entry)()) {
  getcontext(&c.uc);
  c.uc.uc_stack.ss_sp = stack;
  c.uc.uc_stack.ss_size = size;
  c.uc.uc_link = &back.uc;
  makecontext(&c.uc, entry, 0);
}
#endif
