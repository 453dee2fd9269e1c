"""The Julia drop-in files cannot be executed here (no Julia toolchain), so their ccall signatures are checked
mechanically against the C headers they bind: every `ccall((:sym, libadmmtv), Ret, (ArgTypes...), args...)` in
admm_deconv_b200/julia/*.jl must name a function declared in include/*.h with the same number of parameters, matching
parameter kinds (pointer / scalar and base type) and return type, and pass exactly as many values as it declares types.
Also: the four layer types are declared literally with the reference's `Flux.@layer ... trainable=(...)` tuples."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _split_top(s: str):
    """split on commas that are not nested in (), {} or []"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _balanced(s: str, start: int):
    """index just past the parenthesis that closes the one at s[start]"""
    depth = 0
    for i in range(start, len(s)):
        if s[i] == "(":
            depth += 1
        elif s[i] == ")":
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced")


def c_prototypes():
    protos = {}
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        for m in re.finditer(r"\b(int|void|const char\s*\*)\s*(admmtv_[a-z0-9_]+)\s*\(", src):
            end = _balanced(src, m.end() - 1)
            args = src[m.end():end - 1].strip()
            params = [] if args in ("", "void") else _split_top(args)
            protos[m.group(2)] = (m.group(1).replace(" ", ""), [_c_kind(p) for p in params])
    return protos


def _c_kind(p: str):
    p = re.sub(r"\bconst\b", "", p).strip()
    if re.search(r"\[\d*\]\s*$", p):                       # size_t out[4]
        p = re.sub(r"\s*\w+\[\d*\]\s*$", "*", p)
    if "(*" in p:                                            # function pointer
        return ("ptr", "fn")
    ptr = p.count("*")
    base = re.sub(r"[\*\s]+\w*$", "", p).strip() if ptr else re.sub(r"\s+\w+$", "", p).strip()
    base = base.replace("struct ", "")
    return ("ptr" if ptr else "val", base)


JL = {"Cint": ("val", {"int", "int32_t"}), "Cfloat": ("val", {"float"}), "Int64": ("val", {"int64_t"}), "Clonglong": ("val", {"int64_t"}),
      "Csize_t": ("val", {"size_t"})}


def _jl_matches(jt: str, ck) -> bool:
    kind, base = ck
    m = re.fullmatch(r"(Ref|Ptr|CuPtr)\{(\w+)\}", jt)
    if m:
        if kind != "ptr":
            return False
        inner = m.group(2)
        ok = {"Cfloat": {"float"}, "Cvoid": {"void"}, "Csize_t": {"size_t"}, "Desc": {"admmtv_desc"}, "Cdouble": {"double"},
              "UInt8": {"void", "uint8_t"}, "Int64": {"int64_t"}}.get(inner, set())
        return base in ok
    if jt in JL:
        return kind == JL[jt][0] and base in JL[jt][1]
    return False


def jl_ccalls(path):
    src = open(path).read()
    src = re.sub(r"#[^\n]*", "", src)
    out = []
    for m in re.finditer(r"ccall\(", src):
        end = _balanced(src, m.end() - 1)
        parts = _split_top(src[m.end():end - 1])
        sym = re.match(r"\(:(\w+),\s*libadmmtv\)", parts[0]).group(1)
        ret = parts[1]
        types = _split_top(parts[2].strip()[1:-1]) if parts[2].strip() != "()" else []
        types = [t for t in types if t]
        out.append((sym, ret, types, parts[3:]))
    return out


def test_every_julia_ccall_matches_a_header_prototype():
    protos = c_prototypes()
    assert "admmtv_forward" in protos and "admmtv_gmsd_forward" in protos and "admmtv_forward_host" in protos
    seen = set()
    for path in sorted(glob.glob(os.path.join(ROOT, "admm_deconv_b200", "julia", "*.jl"))):
        calls = jl_ccalls(path)
        assert calls, path
        for sym, ret, types, values in calls:
            assert sym in protos, (path, sym, "not declared in include/*.h")
            cret, cparams = protos[sym]
            assert {"Cint": "int", "Cstring": "constchar*", "Cvoid": "void"}[ret] == cret, (sym, ret, cret)
            assert len(types) == len(cparams), (sym, "arity", len(types), len(cparams))
            assert len(values) == len(types), (sym, "passes", len(values), "values for", len(types), "types")
            for i, (jt, ck) in enumerate(zip(types, cparams)):
                assert _jl_matches(jt, ck), (sym, "argument", i, jt, ck)
            seen.add(sym)
    # the drop-in binds the whole forward / backward / host / loss surface
    assert {"admmtv_forward", "admmtv_backward", "admmtv_forward_host", "admmtv_workspace_bytes", "admmtv_strerror",
            "admmtv_gmsd_forward", "admmtv_gmsd_backward", "admmtv_ssim_forward", "admmtv_ssim_backward"} <= seen


def test_julia_layers_are_declared_literally_like_the_reference():
    """deconv_admm.jl:55,107,161,209: Flux.@layer needs the trainable tuple as source text."""
    src = open(os.path.join(ROOT, "admm_deconv_b200", "julia", "ADMMTV.jl")).read()
    want = {"ADMMDeconvF1": "(weight, bias, ρ,)", "ADMMDeconvF2": "(weight, bias, λ,)", "ADMMDeconvF3": "(weight, bias,)",
            "ADMMDeconv": "(weight, bias, λ, ρ,)"}
    for name, tup in want.items():
        assert re.search(r"mutable struct %s\{F,A,N,V,M,B,C,D\}\s+σ::F\s+weight::A\s+bias::V\s+λ::N\s+ρ::M\s+iters::B\s+iso::C\s+creg::D\s+end" % name, src), name
        assert f"Flux.@layer {name} trainable={tup}" in src, name
    assert "$trainables" not in src and "@eval" not in src
    assert "GC.@preserve" in src and "admmtv_forward_host" in src and "FLAG_NOGRAD_REPEAT" in src
