# Generates ic.ptx: loops whose body is N DFMAs (8 independent chains, one constant-bank operand per
# DFMA pair, like the polynomial code) to find the loop-body size at which instruction fetch starts
# to starve the FP64 pipe on B200.  See experiments/README.md for the result.
sizes = [64, 128, 256, 384, 512, 768, 1024, 2048, 4096, 8192]
L = [".version 8.8", ".target sm_100a", ".address_size 64", ".const .align 8 .f64 CC[1024];"]
for N in sizes:
    L.append(".visible .entry body%d(.param .u64 out, .param .u32 iters)" % N)
    L.append(".maxntid 256,1,1\n.minnctapersm 1")
    L.append("{\n.reg .b64 %rd<4>; .reg .b32 %r<8>; .reg .pred %p; .reg .f64 %a<8>, %k<" + str(N) + ">;")
    L.append("ld.param.u64 %rd1,[out]; ld.param.u32 %r1,[iters]; mov.u32 %r2,%tid.x; mov.u32 %r3,%ctaid.x;")
    L.append("cvt.rn.f64.u32 %a0,%r2;")
    for i in range(1, 8):
        L.append("add.f64 %%a%d,%%a0,0d3FF0000000000000;" % i)
    L.append("mov.u32 %r4,0;\nLOOP" + str(N) + ":")
    for j in range(N):
        if j % 2 == 0:
            L.append("ld.const.f64 %%k%d,[CC+%d];" % (j, 8 * ((j // 2 * 7) % 1024)))
        L.append("fma.rn.f64 %%a%d,%%a%d,%%k%d,%%a%d;" % (j % 8, j % 8, j - (j % 2), (j + 3) % 8))
    L.append("add.u32 %r4,%r4,1; setp.lt.u32 %p,%r4,%r1; @%p bra LOOP" + str(N) + ";")
    for i in range(1, 8):
        L.append("add.f64 %%a0,%%a0,%%a%d;" % i)
    L.append("cvta.to.global.u64 %rd2,%rd1; mad.lo.u32 %r5,%r3,256,%r2; mul.wide.u32 %rd3,%r5,8; add.u64 %rd2,%rd2,%rd3; st.global.f64 [%rd2],%a0; ret;\n}")
open("ic.ptx", "w").write("\n".join(L) + "\n")
