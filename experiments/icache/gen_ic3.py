# Probe: compact loops (fit the L0 I-cache) where every DFMA pair takes a NEW warp-uniform coefficient from the
# constant bank through a loop-carried (dynamic) offset -- the access pattern of a constant-bank coefficient tape.
#   mode c64 : ld.const.f64      one LDCU.64 per 2 DFMA
#   mode c128: ld.const.v2.f64   one LDCU.128 per 4 DFMA
# body = DFMAs per loop iteration; the tape pointer wraps inside a 56 KB window.
sizes = [64, 128, 256]
L = [".version 8.8", ".target sm_100a", ".address_size 64", ".const .align 16 .f64 CC[7168];"]
for mode in ("c64", "c128"):
    for N in sizes:
        L.append(".visible .entry %s_body%d(.param .u64 out, .param .u32 iters)" % (mode, N))
        L.append(".maxntid 256,1,1\n.minnctapersm 1")
        L.append("{\n.reg .b64 %rd<6>; .reg .b32 %r<10>; .reg .pred %p, %q; .reg .f64 %a<8>, %k<" + str(N + 2) + ">;")
        L.append("ld.param.u64 %rd1,[out]; ld.param.u32 %r1,[iters]; mov.u32 %r2,%tid.x; mov.u32 %r3,%ctaid.x;")
        L.append("cvt.rn.f64.u32 %a0,%r2;")
        for i in range(1, 8):
            L.append("add.f64 %%a%d,%%a0,0d3FF0000000000000;" % i)
        L.append("mov.u32 %r4,0; mov.u64 %rd4, CC; mov.u32 %r6, 0;")
        L.append("LOOP_%s_%d:" % (mode, N))
        for j in range(N):
            if mode == "c64" and j % 2 == 0:
                L.append("ld.const.f64 %%k%d,[%%rd4+%d];" % (j, 8 * (j // 2)))
            if mode == "c128" and j % 4 == 0:
                L.append("ld.const.v2.f64 {%%k%d,%%k%d},[%%rd4+%d];" % (j, j + 2, 16 * (j // 4)))
            L.append("fma.rn.f64 %%a%d,%%a%d,%%k%d,%%a%d;" % (j % 8, j % 8, j - (j % 2), (j + 3) % 8))
        adv = 8 * (N // 2)
        L.append("add.u32 %%r6,%%r6,%d; setp.ge.u32 %%q,%%r6,%d; @%%q mov.u32 %%r6,0; cvt.u64.u32 %%rd5,%%r6; mov.u64 %%rd4,CC; add.u64 %%rd4,%%rd4,%%rd5;" % (adv, 57344 - adv))
        L.append("add.u32 %%r4,%%r4,1; setp.lt.u32 %%p,%%r4,%%r1; @%%p bra LOOP_%s_%d;" % (mode, N))
        for i in range(1, 8):
            L.append("add.f64 %%a0,%%a0,%%a%d;" % i)
        L.append("cvta.to.global.u64 %rd2,%rd1; mad.lo.u32 %r5,%r3,256,%r2; mul.wide.u32 %rd3,%r5,8; add.u64 %rd2,%rd2,%rd3; st.global.f64 [%rd2],%a0; ret;\n}")
open("ic3.ptx", "w").write("\n".join(L) + "\n")
