#!/bin/bash
# Builds scripts/abl/lib_fwdstats.so: the forward kernel with clock64 phase counters
# (a copy of csrc is patched; the tree is not touched).
set -e
rm -rf /tmp/csrc_stats && cp -r argweaver_b200/csrc /tmp/csrc_stats
python - <<'PY'
import os
p='/tmp/csrc_stats/awb_forward_fast.cuh'
s=open(p).read()
def rep(a,b,cnt=1):
    global s
    assert a in s, a[:60]
    s=s.replace(a,b,cnt)
rep("template <int N> struct AwbInt","__device__ __forceinline__ long long awb_clk() { long long t; asm volatile(\"mov.u64 %0, %%clock64;\" : \"=l\"(t) :: \"memory\"); return t; }\n#define FT_DECL long long ft[10] = {0,0,0,0,0,0,0,0,0,0}; long long ft_last = awb_clk();\n#define FT(i) do { long long t_ = awb_clk(); ft[i] += t_ - ft_last; ft_last = t_; } while (0)\n\ntemplate <int N> struct AwbInt")
rep('''    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory");
}''','''    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory");
    unsigned dummy;
    asm volatile("ld.shared.u32 %0, [%1];\\n\\tmov.u32 %0, %0;" : "=r"(dummy) : "r"(0u) : "memory");
    if (dummy == 0x12345678u) asm volatile("trap;");
}''')
# norm
rep("        int bad_site = -1;\n","        int bad_site = -1;\n        FT_DECL\n")
rep("            const unsigned Fa_s = Fs_s + (site & 1) * RSTR + 8u * lane;\n            awb_bar_sync(2, NB2);","            const unsigned Fa_s = Fs_s + (site & 1) * RSTR + 8u * lane;\n            FT(0);\n            awb_bar_sync(2, NB2);\n            FT(1);")
rep("            if (site == n - 1 && lane == 0) {\n                // (a segment","            if (site == n - 1 && lane == 0 && blockIdx.x == 0)\n                printf(\"norm cycles/site: work %lld bar2wait %lld\\n\", ft[0]/n, ft[1]/n);\n            if (site == n - 1 && lane == 0) {\n                // (a segment")
# scribe
rep("        int site = 0;\n        for (int b = bbeg; b < bend; b++) {\n            const int blen = (b == bextra) ? 1 : blocklensg[b];\n            const int sc_start","        int site = 0;\n        FT_DECL\n        for (int b = bbeg; b < bend; b++) {\n            const int blen = (b == bextra) ? 1 : blocklensg[b];\n            const int sc_start")
rep("                const unsigned Fp_s = Fs_s + (site & 1) * RSTR;\n                awb_bar_sync(1, NB1);","                const unsigned Fp_s = Fs_s + (site & 1) * RSTR;\n                FT(0);\n                awb_bar_sync(1, NB1);\n                FT(1);")
rep("                double v = v0 + v1;","                double v = v0 + v1;\n                if (v == 1.2345e300) printf(\"x\");\n                FT(6);")
rep("                if (sc_last)\n                    awb_sts(Fp_s + 8u * (unsigned) sc_row, v);\n                awb_bar_sync(3, AWB_FWD_FSCRIBES);","                if (v == 1.2345e300) printf(\"x\");\n                FT(7);\n                if (sc_last)\n                    awb_sts(Fp_s + 8u * (unsigned) sc_row, v);\n                FT(2);\n                awb_bar_sync(3, AWB_FWD_FSCRIBES);\n                FT(3);")
rep("                            (ra + rb) + (rc + rd));\n                }\n                awb_bar_sync(2, NB2);","                            (ra + rb) + (rc + rd));\n                }\n                FT(4);\n                awb_bar_sync(2, NB2);\n                FT(5);")
rep("        __syncthreads();                                   // final barrier\n        return;\n    }\n\n    // =====================================================================\n    // compute warps","        if ((sl == 0 || sl == 32) && blockIdx.x == 0)\n            printf(\"scribe %d cycles/site: misc %lld bar1wait %lld sum %lld scan %lld store %lld bar3wait %lld R %lld bar2wait %lld\\n\", sl, ft[0]/n, ft[1]/n, ft[6]/n, ft[7]/n, ft[2]/n, ft[3]/n, ft[4]/n, ft[5]/n);\n        __syncthreads();                                   // final barrier\n        return;\n    }\n\n    // =====================================================================\n    // compute warps")
# scribes: their own work at a block start (loop top to the site loop)
rep("            const int sc_start = sc_startg[(size_t) b * AWB_NSCRIBE + sl];","            const long long tbs = awb_clk();\n            const int sc_start = sc_startg[(size_t) b * AWB_NSCRIBE + sl];")
rep("            for (int i = 0; i < blen; i++, site++) {\n                const unsigned Fp_s","            { const long long t_ = awb_clk(); ft[8] += t_ - tbs; ft_last = t_; }\n            for (int i = 0; i < blen; i++, site++) {\n                const unsigned Fp_s")
rep("scribe %d cycles/site: misc %lld","scribe %d block start %lld cycles/block; cycles/site: misc %lld")
rep("sl, ft[0]/n, ft[1]/n, ft[6]/n","sl, ft[8]/(bend-bbeg), ft[0]/n, ft[1]/n, ft[6]/n")
# compute
rep("    // ---- my state in the current block\n","    // ---- my state in the current block\n    FT_DECL\n    long long tb0 = 0, tbound = 0; bool fs = false;\n    int posb = 0; long long tprev = awb_clk(); long long tp[10] = {0,0,0,0,0,0,0,0,0,0}; int np[10] = {0,0,0,0,0,0,0,0,0,0};\n")
rep("        awb_sts(zaddr, c);\n        awb_bar_sync(1, NB1);\n\n        // branch scans","        awb_sts(zaddr, c);\n        FT(0);\n        awb_bar_sync(1, NB1);\n        FT(1);\n        if (fs) { tbound += ft_last - tb0; fs = false; }\n\n        // branch scans")
rep("            e = (kd == AWB_SITE_VARIANT) ? (live ? *nxt : inv_e) : em;\n        awb_bar_sync(2, NB2);","            e = (kd == AWB_SITE_VARIANT) ? (live ? *nxt : inv_e) : em;\n        FT(2);\n        awb_bar_sync(2, NB2);\n        FT(3);\n        { const int pi = posb < 9 ? posb : 9; tp[pi] += ft_last - tprev; np[pi]++; tprev = ft_last; posb++; }")
rep("            awb_bar_sync(2, NB2);\n            if (b == bend - 1)\n                break;\n","            awb_bar_sync(2, NB2);\n            FT(6);\n            tb0 = ft_last; fs = true;\n            tprev = ft_last; posb = 0;\n            if (b == bend - 1)\n                break;\n")
if os.environ.get("FWD_LCDETAIL"):
    rep("        active = (tj != 0xFFFF);\n","        active = (tj != 0xFFFF);\n        if (bb != bbeg) { if (tj == 0xFFF1) printf(\"x\"); FT(7); }\n")
    rep("        zaddr = active ? zT_s + 8u * (unsigned) tpos : dummy_s;\n","        if (bb != bbeg) { if (tpos + atime + cage + node == -12345 || inv_e == 1.2345e300) printf(\"x\"); FT(8); }\n        zaddr = active ? zT_s + 8u * (unsigned) tpos : dummy_s;\n")
    rep("        if (live) {\n            const double *lin = ling + (size_t) bb * 7 * T;","        if (bb != bbeg) { if (nl == 77) printf(\"x\"); FT(9); }\n        if (live) {\n            const double *lin = ling + (size_t) bb * 7 * T;")
rep("            load_compute(b + 1);\n            double sum = 0.0;","            load_compute(b + 1);\n            FT(4);\n            double sum = 0.0;")
rep("            c = sum * e * scale;\n","            c = sum * e * scale;\n            FT(5);\n")
rep("    // ---- the last two columns: their 1/norm is complete after the final barrier\n    __syncthreads();","    if ((tid == 0 || tid == NS - 32) && blockIdx.x == 0)\n        printf(\"compute tid %d cycles/site: A-phase %lld bar1wait %lld B-phase %lld bar2wait %lld | per block (%d blocks): last site %lld load_compute %lld gather %lld, bar2(last site) -> bar1(first site) %lld\\n\", tid, ft[0]/n, ft[1]/n, ft[2]/n, ft[3]/n, bend - bbeg, ft[6]/(bend-bbeg), ft[4]/(bend-bbeg), ft[5]/(bend-bbeg), tbound/(bend-bbeg)); if (tid == 0 && blockIdx.x == 0) { printf(\"   cycles bar2->bar2 by position of the site in its block:\"); for (int i = 0; i < 10; i++) printf(\" [%d] %lld (n=%d)\", i, np[i] ? tp[i] / np[i] : 0, np[i]); printf(\"\\n\"); }\n    if ((tid == 0 || tid == NS - 32) && blockIdx.x == 0) printf(\"   load_compute detail tid %d: tmap %lld per-state %lld match %lld (lin = load_compute above)\\n\", tid, ft[7]/(bend-bbeg), ft[8]/(bend-bbeg), ft[9]/(bend-bbeg));\n    // ---- the last two columns: their 1/norm is complete after the final barrier\n    __syncthreads();")
if os.environ.get("FWD_SUMSRC"):
    # scribes sum from a region nobody writes in the site loop (timing experiment only)
    s=s.replace("const unsigned z_s = zT_s + 8u * (unsigned) sc_start;","const unsigned z_s = col_s + 8u * (unsigned) (sc_start % 64);")
if os.environ.get("FWD_NOSCAN"):
    import re
    s=re.sub(r"site_step\(AwbInt<[^;]*>\(\)\);", "site_step(AwbInt<0>());", s)
open(p,"w").write(s)
PY
mkdir -p scripts/abl
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared --fmad=true -I include -I /tmp/csrc_stats -o scripts/abl/lib_fwdstats.so /tmp/csrc_stats/awb_api.cu /tmp/csrc_stats/awb_compat.cu -lcudart 2>&1 | grep -i "error" || true
