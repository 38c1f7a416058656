# Variants of the loop-size probe where the per-pair coefficient comes from
#   mode "lds"  : ld.shared.f64 (broadcast)          one load per 2 DFMA
#   mode "lds2" : ld.shared.v2.f64                   one load per 4 DFMA
#   mode "ldg2" : ld.global.nc.v2.f64 (L1-resident)  one load per 4 DFMA
import sys
sizes = [256, 1024, 4096, 8192]
L = [".version 8.8", ".target sm_100a", ".address_size 64"]
for mode in ("lds", "lds2", "ldg2"):
  for N in sizes:
    L.append(".visible .entry %s_body%d(.param .u64 out, .param .u32 iters, .param .u64 tab)" % (mode, N))
    L.append(".maxntid 256,1,1\n.minnctapersm 1")
    L.append("{\n.reg .b64 %rd<8>; .reg .b32 %r<10>; .reg .pred %p; .reg .f64 %a<8>, %k<" + str(N + 2) + ">;")
    L.append(".shared .align 16 .f64 SC[1024];")
    L.append("ld.param.u64 %rd1,[out]; ld.param.u32 %r1,[iters]; ld.param.u64 %rd4,[tab]; cvta.to.global.u64 %rd5,%rd4; mov.u32 %r2,%tid.x; mov.u32 %r3,%ctaid.x;")
    # fill shared table: 1024 doubles, 256 threads x 4
    for q in range(4):
        L.append("mad.lo.u32 %%r6,%%r2,4,%d; mul.wide.u32 %%rd6,%%r6,8; add.u64 %%rd7,%%rd5,%%rd6; ld.global.f64 %%k0,[%%rd7]; mov.u32 %%r7,SC; shl.b32 %%r8,%%r6,3; add.u32 %%r7,%%r7,%%r8; st.shared.f64 [%%r7],%%k0;" % q)
    L.append("bar.sync 0;")
    L.append("cvt.rn.f64.u32 %a0,%r2;")
    for i in range(1, 8):
        L.append("add.f64 %%a%d,%%a0,0d3FF0000000000000;" % i)
    L.append("mov.u32 %%r4,0;\nLOOP_%s_%d:" % (mode, N))
    for j in range(N):
        if mode == "lds":
            if j % 2 == 0:
                L.append("ld.shared.f64 %%k%d,[SC+%d];" % (j, 8 * ((j // 2 * 7) % 1024)))
            kk = j - (j % 2)
        elif mode == "lds2":
            if j % 4 == 0:
                L.append("ld.shared.v2.f64 {%%k%d,%%k%d},[SC+%d];" % (j, j + 2, 16 * ((j // 4 * 7) % 512)))
            kk = j - (j % 2)
        else:
            if j % 4 == 0:
                L.append("ld.global.nc.v2.f64 {%%k%d,%%k%d},[%%rd5+%d];" % (j, j + 2, 16 * ((j // 4 * 7) % 512)))
            kk = j - (j % 2)
        L.append("fma.rn.f64 %%a%d,%%a%d,%%k%d,%%a%d;" % (j % 8, j % 8, kk, (j + 3) % 8))
    L.append("add.u32 %%r4,%%r4,1; setp.lt.u32 %%p,%%r4,%%r1; @%%p bra LOOP_%s_%d;" % (mode, N))
    for i in range(1, 8):
        L.append("add.f64 %%a0,%%a0,%%a%d;" % i)
    L.append("cvta.to.global.u64 %rd2,%rd1; mad.lo.u32 %r5,%r3,256,%r2; mul.wide.u32 %rd3,%r5,8; add.u64 %rd2,%rd2,%rd3; st.global.f64 [%rd2],%a0; ret;\n}")
open("ic2.ptx", "w").write("\n".join(L) + "\n")
