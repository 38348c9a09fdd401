"""Static check of the programmatic-dependent-launch discipline (DESIGN.md "launch chain").

A kernel launched with the PDL attribute may start while the kernel before it in the stream is still running; it
is only correct if every thread executes griddepcontrol.wait (pdl_wait()) before it touches global memory -- and
that holds transitively only if EVERY kernel of the chain does so on every path (or, instead, acquires a flag / word
that the previous kernel's last block released once its outputs were written: chain_wait(), the peer flags).  The CPU emulation cannot see a
violation (its launches run one after the other), a GPU run only sometimes: round 2 shipped k_paint_strips without
the wait for one GPU run and got stale strip tables.  So the rule is checked on the sources: every kernel that
ddc_api.cu launches with a `pdl` argument that can be true calls pdl_wait() before its first `if (...) return`.
"""
import os
import re

from conftest import ROOT

API = os.path.join(ROOT, "domain_decomp_b200", "csrc", "ddc_api.cu")
KERNELS = os.path.join(ROOT, "domain_decomp_b200", "csrc", "ddc_kernels.cuh")


def _launches():
    src = open(API).read()
    out = []
    for m in re.finditer(r"launch_k\(", src):
        depth, i = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        args, cur, d = [], "", 0
        for ch in src[m.end():i - 1]:
            d += {"(": 1, "<": 0, ")": -1}.get(ch, 0)
            if ch == "," and d == 0:
                args.append(cur.strip())
                cur = ""
            else:
                cur += ch
        out.append(args)
    return out


def _kernel_body(name):
    src = re.sub(r"/\*.*?\*/", lambda c: " " * len(c.group(0)), open(KERNELS).read(), flags=re.S)  # braces in comments
    src = re.sub(r"//[^\n]*", lambda c: " " * len(c.group(0)), src)
    m = re.search(r"__global__[^{;]*\b%s\s*\(" % re.escape(name), src)
    assert m, "kernel %s not found" % name
    i = src.index("{", m.end())
    depth, j = 1, i + 1
    while depth:
        depth += {"{": 1, "}": -1}.get(src[j], 0)
        j += 1
    return src[i:j]


def test_every_pdl_launched_kernel_waits_first():
    seen = set()
    for args in _launches():
        if len(args) < 6:  # the definition of launch_k itself
            continue
        kernel_expr, pdl = args[0], args[5]
        if pdl == "false":
            continue
        for name in set(re.findall(r"\bk_\w+", kernel_expr)):
            seen.add(name)
            body = _kernel_body(name)
            # chain_wait(word) polls the word the previous kernel's last block publishes and falls back to pdl_wait()
            # ("flags instead of kernel boundaries")
            waits = [body.find(w) for w in ("pdl_wait()", "chain_wait(") if w in body]
            assert waits, "%s is launched with PDL but never calls pdl_wait() / chain_wait()" % name
            w = min(waits)
            r = re.search(r"\breturn\b", body)
            assert r is None or w < r.start(), "%s can return before pdl_wait(): breaks the transitive completion" % name
            # nothing of global memory before the wait: no pointer dereference / index into a kernel parameter
            head = body[:w].replace("plan->ts[TS_RES", "")  # (DDC_DEBUG_TS stamps "block is on an SM": write-only, diagnostic)
            assert "plan->" not in head and "sc->" not in head, "%s reads device state before pdl_wait()" % name
    assert {"k_scan_mask", "k_xcuts", "k_strip_rows_scan", "k_ycuts", "k_paint_strips", "k_sum_cols"} <= seen, seen
