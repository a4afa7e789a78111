#!/usr/bin/env python
"""Extract the Python-visible signatures of the reference's 13 `omniserve_backend` extension modules by PARSING the
reference's own sources (the `m.def("name", &fn ...)` registrations and the C++ parameter lists of `fn`), and write
them to tests/golden/ref_api.json.  The CPU test tests/test_host_logic.py compares the ctypes mirror
(omniserve_b200/backend/*.py) against this file -- arity AND parameter order -- and, when /root/reference is present,
re-runs this parser and requires the committed JSON to be up to date.

    python tests/golden/make_ref_api.py [--check]

TEST INFRASTRUCTURE: reads /root/reference (read-only); nothing from it is copied except identifier names.
"""
from __future__ import annotations

import json
import os
import re
import sys

REF = os.environ.get("OMNISERVE_REFERENCE", "/root/reference")
CSRC = os.path.join(REF, "kernels", "csrc")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_api.json")
FA = "fused_attention"

# module -> directories / files that hold its registration and the registered functions (kernels/setup.py:156-333)
MODULE_SOURCES = {
    "qgemm_w4a8_per_chn": ["qgemm/w4a8_per_chn"],
    "qgemm_w4a8_per_group": ["qgemm/w4a8_per_group"],
    "qgemm_w8a8": ["qgemm/w8a8"],
    "fused_kernels": ["fused.cpp", "fused_kernels.cu"],
    "layernorm_ops": ["layernorm.cpp", "layernorm_kernels.cu"],
    "activation_ops": ["activation.cpp", "activation_kernels.cu"],
    "fused_attention_pure_dense": [f"{FA}/fused_attention_pure_dense"],
    "fused_attention_fine_grained_dense": [f"{FA}/fused_attention_fine_grained/dense_attention",
                                           f"{FA}/fused_attention_fine_grained/fine_grained_common", f"{FA}/common"],
    "fused_attention_fine_grained_sparse": [f"{FA}/fused_attention_fine_grained/sparse_attention",
                                            f"{FA}/fused_attention_fine_grained/fine_grained_common", f"{FA}/common"],
    "fused_attention_per_tensor_dense": [f"{FA}/fused_attention_per_tensor/dense_attention",
                                         f"{FA}/fused_attention_per_tensor/per_tensor_common", f"{FA}/common"],
    "fused_attention_per_tensor_sparse": [f"{FA}/fused_attention_per_tensor/sparse_attention",
                                          f"{FA}/fused_attention_per_tensor/per_tensor_common", f"{FA}/common"],
    "fused_attention_selector": [f"{FA}/sparse_utils/KVPageSelector"],
    "fused_attention_ctx_pool": [f"{FA}/sparse_utils/ContextPool"],
}


def _strip_comments(s: str) -> str:
    s = re.sub(r"/\*.*?\*/", " ", s, flags=re.S)
    return re.sub(r"//[^\n]*", " ", s)


def _files(entries):
    out = []
    for e in entries:
        p = os.path.join(CSRC, e)
        if os.path.isdir(p):
            for f in sorted(os.listdir(p)):
                if f.endswith((".cpp", ".cu", ".h", ".cuh", ".hpp")):
                    out.append(os.path.join(p, f))
        elif os.path.exists(p):
            out.append(p)
    return out


def _balanced(s: str, i: int) -> str:
    """s[i] == '(' -> the text between it and its matching ')'."""
    depth, j = 0, i
    while j < len(s):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return s[i + 1:j]
        j += 1
    raise ValueError("unbalanced")


def _params(arglist: str):
    parts, depth, cur = [], 0, ""
    for ch in arglist:
        if ch in "<([{":
            depth += 1
        elif ch in ">)]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur)
    names = []
    for p in parts:
        p = p.split("=")[0].strip()
        m = re.search(r"([A-Za-z_]\w*)\s*$", p)
        names.append(m.group(1) if m else "?")
    return names


def _find_function(texts, fn):
    """Parameter names of the free function `fn` (first definition / declaration with a parameter list found)."""
    best = None
    for txt in texts:
        for m in re.finditer(r"(?<![\w&.:>])" + re.escape(fn) + r"\s*\(", txt):
            head = txt[max(0, m.start() - 80):m.start()]
            if not re.search(r"[\w>&*]\s*$", head) or re.search(r"(return|=|\(|,)\s*$", head):
                continue  # a call, not a declaration
            try:
                args = _balanced(txt, m.end() - 1)
            except ValueError:
                continue
            names = _params(args)
            if names and all(n != "?" for n in names) and (best is None or len(names) > len(best)):
                best = names
    return best


def extract():
    api = {}
    for mod, entries in MODULE_SOURCES.items():
        texts = []
        for f in _files(entries):
            try:
                texts.append(_strip_comments(open(f, errors="ignore").read()))
            except OSError:
                pass
        funcs = {}
        for txt in texts:
            if "PYBIND11_MODULE" not in txt:
                continue
            body = txt[txt.index("PYBIND11_MODULE"):]
            for m in re.finditer(r"m\.def\(\s*\"(\w+)\"\s*,", body):
                name = m.group(1)
                rest = body[m.end():m.end() + 400]
                ov = re.match(r"\s*py::overload_cast<(.*?)>\(\s*&\s*(\w+)\s*\)", rest, flags=re.S)
                if ov:
                    # overloads: record each by its C++ type list length (the Python name is shared)
                    n = len(_params(ov.group(1)))
                    funcs.setdefault(name, {"overload_arities": []})
                    funcs[name].setdefault("overload_arities", []).append(n)
                    target = ov.group(2)
                else:
                    t = re.match(r"\s*&?\s*(\w+)", rest)
                    target = t.group(1) if t else None
                if not target:
                    continue
                names = _find_function(texts, target)
                if names:
                    ent = funcs.setdefault(name, {})
                    if "params" not in ent or len(names) > len(ent["params"]):
                        ent["params"] = names
                # py::arg names, when the registration spells them (layernorm.cpp / activation.cpp)
                reg = body[m.start():]
                try:
                    reg = _balanced(reg, reg.index("("))
                except ValueError:
                    reg = ""
                kw = re.findall(r"py::arg\(\s*\"(\w+)\"\s*\)", reg)
                if kw:
                    funcs[name]["py_args"] = kw
        if funcs:
            api[mod] = funcs
    return api


def main():
    if not os.path.isdir(CSRC):
        print(f"{CSRC} not present: nothing to do")
        return 0
    api = extract()
    txt = json.dumps(api, indent=1, sort_keys=True) + "\n"
    if "--check" in sys.argv:
        return 0 if os.path.exists(OUT) and open(OUT).read() == txt else 1
    open(OUT, "w").write(txt)
    n = sum(len(v) for v in api.values())
    print(f"wrote {OUT}: {len(api)} modules, {n} functions")
    return 0


if __name__ == "__main__":
    sys.exit(main())
