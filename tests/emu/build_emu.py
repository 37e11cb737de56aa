"""Builds tests/emu/_build/libunivs_emu.so: the plain CUDA kernels of univs_b200/csrc compiled by g++ for CPU execution
(tests/emu/cuda_emu.h).  The only source transformation is the launch syntax:
    kernel<<<grid, block, smem, stream>>>(args);   ->   ::emu::launch(dim3(grid), dim3(block), [=]() { kernel(args); });
Everything else -- kernels, argument validation, the extern "C" entry points -- is compiled as written."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "univs_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
SOURCES = ["common.cu", "groupnorm.cu", "swin_glue.cu", "decoder_glue.cu", "elementwise.cu", "msda.cu",
           "swin_window_attn_tc.cu", "swin_window_attn_tc2.cu", "mha_tc.cu", "mask_einsum_mc.cu", "mask_einsum_tc.cu", "gemm_tc.cu",
           "swin_window_attn.cu", "mha.cu", "mask_einsum.cu"]          # the last three: mma.sync / cp.async kernels
HEADERS = ["common.cuh", "rowwise.cuh", "tc05_math.cuh"]      # tc05.cuh itself is replaced by tests/emu/tc05.cuh
CUDA_INCLUDE = os.environ.get("CUDA_INCLUDE", "/usr/local/cuda/include")


def _matching(text, start, open_ch, close_ch):
    """index just past the bracket that closes the one at text[start]"""
    depth = 0
    for i in range(start, len(text)):
        if text[i] == open_ch:
            depth += 1
        elif text[i] == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced brackets")


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{<" and not (ch == "<" and depth == 0 and False):
            depth += ch in "([{"
        if ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite_launches(text):
    out, pos = "", 0
    while True:
        i = text.find("<<<", pos)
        if i < 0:
            return out + text[pos:]
        # kernel expression: identifier, optionally with template arguments, right before <<<
        j = i
        if text[j - 1] == ">":                     # name<...>
            depth = 0
            while True:
                j -= 1
                if text[j] == ">":
                    depth += 1
                elif text[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
        while j > 0 and (text[j - 1].isalnum() or text[j - 1] in "_:"):
            j -= 1
        kernel = text[j:i]
        k = text.find(">>>", i)
        cfg = _split_top(text[i + 3:k])
        a0 = text.index("(", k)
        a1 = _matching(text, a0, "(", ")")
        args = text[a0 + 1:a1 - 1]
        out += text[pos:j] + f"::emu::launch(dim3({cfg[0]}), dim3({cfg[1]}), [=]() {{ {kernel}({args}); }})"
        pos = a1


def rewrite_launch_ex(text):
    """cudaLaunchKernelEx(&cfg, kernel, args...)  ->  ::emu::launch_ex(cfg, [=]() { kernel(args...); })"""
    out, pos = "", 0
    while True:
        i = text.find("cudaLaunchKernelEx(", pos)
        if i < 0:
            return out + text[pos:]
        a0 = text.index("(", i)
        a1 = _matching(text, a0, "(", ")")
        parts = _split_top(text[a0 + 1:a1 - 1])
        cfg, kernel, args = parts[0].lstrip("&"), parts[1], ", ".join(parts[2:])
        out += text[pos:i] + f"::emu::launch_ex({cfg}, [=]() {{ {kernel}({args}); }})"
        pos = a1


def build(force=False):
    os.makedirs(BUILD, exist_ok=True)
    lib = os.path.join(BUILD, "libunivs_emu.so")
    inputs = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(HERE, f) for f in ("cuda_emu.h", "cuda_emu.cpp", "tc05.cuh", "build_emu.py")]
    if not force and os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(p) for p in inputs):
        return lib
    gen = []
    for f in HEADERS:       # headers are used as they are (UNIVS_CPU_EMU selects host versions of the two PTX helpers)
        with open(os.path.join(CSRC, f)) as src, open(os.path.join(BUILD, f), "w") as dst:
            dst.write(src.read().replace('#include "../../include/univs_b200.h"', f'#include "{ROOT}/include/univs_b200.h"'))
    for f in SOURCES:
        if False:
            pass
        elif f == "mask_einsum_tc.cu":
            # the kernel that IS validated on the B200 keeps private copies of its PTX wrappers; for the emulator they are
            # swapped for tests/emu/tc05.cuh (same names, two of them with another spelling) -- the calibration case of the
            # emulated TMA / descriptor / TMEM semantics.  Any pattern that no longer matches fails the build.
            whole = open(os.path.join(CSRC, f)).read()
            w0 = whole.index("__device__ __forceinline__ uint32_t smem_u32(const void* p)")
            w1 = whole.index("// F16X3 == false: fp32 operands consumed as TF32")
            shim = ("using namespace tc;\n"
                    "__device__ __forceinline__ void tc_fence_before() {}\n__device__ __forceinline__ void tc_fence_after() {}\n"
                    "__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {\n"
                    "  tc::tma_load_3d(smem_u32(dst), map, bar, c0, c1, c2);\n}\n"
                    "#define TMEM_LD_32x32b_X32(taddr, r) UNIVS_TMEM_LD_X32(taddr, r)\n")
            raw = whole[:w0] + shim + whole[w1:]
            raw = raw.replace('#include "common.cuh"', '#include "tc05.cuh"', 1)
            for old, new in [
                ('asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");', "mbar_init_fence();"),
                ('asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));',
                 "tmem_alloc(tmem_slot, 512);"),
                ('asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");', ""),
                ('asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");', "tmem_wait_ld();"),
                ('asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));', "tmem_dealloc(tmem_base, 512);"),
            ]:
                assert raw.count(old) == 1, old
                raw = raw.replace(old, new)
            assert "asm" not in raw.replace("namespace", ""), "unexpected inline assembly left in mask_einsum_tc.cu"
        else:
            raw = open(os.path.join(CSRC, f)).read()
            # the two private m16n8k16 wrappers (inline PTX) -> the emulated warp MMA
            raw = re.sub(r'(void (?:mma_f16|mha_mma_f16)\(float \(&c\)\[4\], const uint32_t \(&a\)\[4\], uint32_t b0, uint32_t b1\) \{)\s*asm volatile\(.*?\);\s*\}',
                         r'\1 ::emu::mma_m16n8k16_f16(c, a, b0, b1); }', raw, flags=re.S)
            assert not re.search(r'asm(?: volatile)?\(\s*"[^"]', raw), f"inline assembly left in {f}"
        if True:
            text = rewrite_launch_ex(rewrite_launches(raw))
            # dynamic shared memory: one emulated buffer per CTA
            text = re.sub(r"extern __shared__ __align__\(\d+\) (unsigned char|float) (\w+)\[\];",
                          r"\1* \2 = reinterpret_cast<\1*>(::emu::dyn_smem());", text)
        path = os.path.join(BUILD, f.replace(".cu", "_emu.cpp"))
        with open(path, "w") as dst:
            dst.write(text)
        gen.append(path)
    cmd = ["g++", "-std=c++20", "-O1", "-w", "-fPIC", "-shared", "-pthread", "-Wl,-Bsymbolic", "-DUNIVS_CPU_EMU", f"-I{CUDA_INCLUDE}", f"-I{HERE}", f"-I{BUILD}", "-include", os.path.join(HERE, "cuda_emu.h"), os.path.join(HERE, "cuda_emu.cpp"), *gen, "-o", lib]
    subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
