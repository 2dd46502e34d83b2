"""CPU: source-level check of the programmatic-dependent-launch rules of gw-depth_b200/csrc (gwd_common.cuh, "PDL").

A kernel launched through gwd_launch() carries cudaLaunchAttributeProgrammaticStreamSerialization, so its CTAs may start while the
kernel in front of it is still running.  That is only correct when
  (1) the kernel executes gwd_pdl_wait() (griddepcontrol.wait) and touches no global memory before it -- checked here as: nothing but
      shared-memory / TMEM / barrier set-up in front of the wait (no __ldg, no dereference of a kernel pointer argument, no TMA);
  (2) gwd_pdl_trigger() (griddepcontrol.launch_dependents) never comes before the kernel's own wait;
  (3) in a kernel that allocates tensor memory the trigger comes after the tcgen05.alloc (a dependent that got the columns first
      would wait for this kernel while holding them: deadlock).
The GPU suite runs with the attribute on (default), so these rules are also exercised by every parity test on the B200."""
import glob
import os
import re

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gw-depth_b200", "csrc")


def _kernels(src):
    """{name: body} of every __global__ function of a translation unit"""
    out = {}
    for m in re.finditer(r"__global__[^;{]*?\b(gwd_\w+)\s*\(", src):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        j = src.index("{", i)
        assert src[i:j].strip() == "", (m.group(1), src[i:j])
        k, depth = j + 1, 1
        while depth:
            depth += {"{": 1, "}": -1}.get(src[k], 0)
            k += 1
        out[m.group(1)] = src[j + 1:k - 1]
    return out


def _sources():
    return {os.path.basename(p): open(p).read() for p in sorted(glob.glob(os.path.join(CSRC, "*.cu")))}


def test_every_programmatically_launched_kernel_waits_before_global_memory():
    launched = 0
    for fname, src in _sources().items():
        kernels = _kernels(src)
        names = set(re.findall(r"gwd_launch\(\s*(gwd_\w+)", src))
        # launches through a local alias (`auto kfn = gwd_tapgemm_kernel<...>`, the `kern` lambda argument, the KERNEL macro argument)
        names |= {k for k in re.findall(r"auto kfn = (gwd_\w+)", src)}
        names |= {k for k in re.findall(r"GWD_CUDA\(launch\((gwd_\w+)<", src)}
        names |= {k for k in re.findall(r"GWD_ROW_DISPATCH\((gwd_\w+)", src)}
        for name in sorted(names):
            assert name in kernels, (fname, name)
            body = kernels[name]
            assert "gwd_pdl_wait()" in body, "%s: %s is launched programmatically but never waits" % (fname, name)
            head = body[:body.index("gwd_pdl_wait()")]
            head = re.sub(r"//[^\n]*", "", head)
            for bad in ("__ldg", "tma_load", "tma_store", "cp.async", "ld.global", "st.global", "atomicAdd", "red."):
                assert bad not in head, "%s: %s touches global memory (%s) ahead of gwd_pdl_wait()" % (fname, name, bad)
            launched += 1
    assert launched >= 17, launched      # the tcgen05 GEMM, both DETR attention kernels, the tcgen05 weight gradient + 13 row kernels


def test_trigger_follows_the_wait_and_the_tmem_allocation():
    seen = 0
    for fname, src in _sources().items():
        for name, body in _kernels(src).items():
            if "gwd_pdl_trigger()" not in body:
                continue
            seen += 1
            t = body.index("gwd_pdl_trigger()")
            assert body.count("gwd_pdl_trigger()") == 1, (fname, name)
            if "gwd_pdl_wait()" in body:
                assert body.index("gwd_pdl_wait()") < t, "%s: %s triggers its dependents before its own wait" % (fname, name)
            if "tcgen05.alloc" in body:
                assert body.rindex("tcgen05.alloc") < t, "%s: %s triggers before it holds its TMEM columns" % (fname, name)
                # ... and after the CTA-wide barrier that follows the allocation (every warp sees the columns as taken)
                between = body[body.rindex("tcgen05.alloc"):t]
                assert "__syncthreads()" in between or "cluster_sync_all()" in between, (fname, name)
    assert seen >= 20, seen


def test_launch_helper_sets_the_attribute_only_when_enabled():
    common = open(os.path.join(CSRC, "gwd_common.cuh")).read()
    assert "cudaLaunchAttributeProgrammaticStreamSerialization" in common and "gwd_pdl_enabled()" in common
    assert re.search(r'getenv\("GWD_PDL"\)', common)
    assert 'griddepcontrol.wait;' in common and 'griddepcontrol.launch_dependents;' in common
