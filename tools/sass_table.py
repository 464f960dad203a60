#!/usr/bin/env python
"""Opcode histogram per kernel of the built library (cuobjdump -sass): which kernels carry tcgen05 (UTCHMMA), TMA
(UTMALDG / UTMASTG), TMEM loads / stores (LDTM / STTM), warp-level MMA (HMMA), packed fp32x2 math (FFMA2 / FADD2 /
FMUL2) and MUFU.   python tools/sass_table.py [lib.so] [name-regex] > profiles/<round>_sass_histograms.md"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "easydgl_b200/csrc/libeasydgl_b200.so"
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else r"attention_f16|attention_tc|attention_mma_kernel<16, 16, 13>|gemm_tc_kernel|gemm_f16_kernel|embed_kernel|topk_kernel|topk_select|time_attention|ln_finalize")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
dem = {}
fn = None; hist = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); hist[fn] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn: hist[fn][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
KEYS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "HMMA", "LDSM", "FFMA2", "FADD2", "FMUL2", "FFMA", "MUFU", "LDL", "STL"]
print("| kernel | instructions | " + " | ".join(KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
for mangled, name in zip(hist, names):
    if not pat.search(name): continue
    h = hist[mangled]
    short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "")).replace("void ", "").replace("edgl::", "")
    print("| `%s` | %d | " % (short[:70], sum(h.values())) + " | ".join(str(h.get(k, 0)) for k in KEYS) + " |")
