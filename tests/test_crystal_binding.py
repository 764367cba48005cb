"""The Crystal `lib LibPhGpu` binding (crystal/src/device/lib_ph_gpu.cr) cannot be compiled here (no
Crystal toolchain), so it is checked mechanically against include/ph_gpu.h: every C function bound
under the same name with the same number of parameters and compatible parameter kinds, every enum
value equal, the descriptor struct laid out alike, and every `LibPhGpu.ph_*` call in the Crystal
sources refers to a bound function with the right number of arguments."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CR_DIR = os.path.join(ROOT, "crystal", "src", "device")


def _c_functions():
    text = open(os.path.join(ROOT, "include", "ph_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for ret, name, params in re.findall(r"\n\s*([\w\s\*]+?)\s*\b(ph_\w+)\s*\(([^;{]*?)\)\s*;", text):
        params = params.strip()
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        out[name] = (ret.strip(), plist)
    return out


def _cr_functions():
    text = open(os.path.join(CR_DIR, "lib_ph_gpu.cr")).read()
    text = re.sub(r"#.*", "", text)
    out = {}
    for name, params, ret in re.findall(r"fun\s+(ph_\w+)\s*(?:\(([^)]*)\))?\s*:\s*([\w:\*]+)", text):
        plist = [p.strip() for p in params.split(",")] if params.strip() else []
        out[name] = (ret, plist)
    return out


def _kind(c_param: str) -> str:
    if "*" in c_param:
        return "ptr"
    if "size_t" in c_param:
        return "size"
    if "int64_t" in c_param:
        return "i64"
    return "i32"


def _cr_kind(cr_param: str) -> str:
    ty = cr_param.split(":", 1)[1].strip()
    if ty.endswith("*"):
        return "ptr"
    return {"LibC::SizeT": "size", "Int64": "i64", "UInt64": "i64", "Int32": "i32"}[ty]


def test_every_c_entry_is_bound_with_the_same_signature_shape():
    c, cr = _c_functions(), _cr_functions()
    assert len(c) >= 45
    assert set(c) == set(cr), (sorted(set(c) - set(cr)), sorted(set(cr) - set(c)))
    for name, (ret, params) in c.items():
        cr_ret, cr_params = cr[name]
        assert len(params) == len(cr_params), name
        assert [_kind(p) for p in params] == [_cr_kind(p) for p in cr_params], name
        want_ret = "ptr" if "*" in ret else ("i64" if "int64_t" in ret else "i32")
        got_ret = "ptr" if cr_ret.endswith("*") else {"Int64": "i64", "Int32": "i32"}[cr_ret]
        assert want_ret == got_ret, name


def test_enum_values_agree_with_the_header():
    import ph_core_b200 as ph
    text = open(os.path.join(CR_DIR, "lib_ph_gpu.cr")).read()
    enums = {m.group(1): dict((k, int(v)) for k, v in re.findall(r"^\s*(\w+)\s*=\s*(\d+)", m.group(2), flags=re.M))
             for m in re.finditer(r"enum (\w+) : Int32\n(.*?)\n  end", text, flags=re.S)}
    prefix = {"DType": "PH_", "Op": "PH_", "Cmp": "PH_", "Unary": "PH_", "Red": "PH_", "HeatMode": "PH_HEAT_", "Status": "PH_"}
    rename = {"FloorDiv": "FLOORDIV", "ArgMax": "ARGMAX", "ArgMin": "ARGMIN", "Example1D": "EXAMPLE1D",
              "ErrCuda": "ERR_CUDA", "ErrInvalid": "ERR_INVALID", "ErrUnsupported": "ERR_UNSUPPORTED", "ErrNccl": "ERR_NCCL",
              "ErrNotInit": "ERR_NOT_INIT", "WAdd": "WADD", "WSub": "WSUB", "WMul": "WMUL", "WPow": "WPOW"}
    checked = 0
    for ename, members in enums.items():
        for member, value in members.items():
            cname = prefix[ename] + rename.get(member, member.upper())
            assert ph.K[cname] == value, (ename, member)
            checked += 1
    assert checked >= 45
    flags = dict(re.findall(r"(FLAG_\w+)\s*=\s*(\d+)_u32", text))
    assert {k: int(v) for k, v in flags.items()} == {k[3:]: ph.K[k] for k in ("PH_FLAG_OVERFLOW", "PH_FLAG_DIV0", "PH_FLAG_NAN", "PH_FLAG_ARGUMENT")}
    assert "MAX_RANK = 8" in text and ph.K.get("PH_MAX_RANK", 8) == 8
    struct = re.search(r"struct Desc\n(.*?)\n  end", text, flags=re.S).group(1)
    fields = [l.split(":")[0].strip() for l in struct.strip().splitlines()]
    assert fields == [f[0] for f in ph.PhDesc._fields_]                       # same order as struct ph_desc


def _call_arity(args: str) -> int:
    depth, n, seen = 0, 0, False
    for ch in args:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        elif ch == "," and depth == 0:
            n += 1
        if not ch.isspace():
            seen = True
    return n + 1 if seen else 0


def test_every_call_site_in_the_crystal_sources_matches_the_binding():
    cr = _cr_functions()
    calls = 0
    for path in glob.glob(os.path.join(CR_DIR, "*.cr")):
        if path.endswith("lib_ph_gpu.cr"):
            continue
        text = open(path).read()
        for m in re.finditer(r"LibPhGpu\.(ph_\w+)", text):
            name = m.group(1)
            assert name in cr, (os.path.basename(path), name)
            rest = text[m.end():]
            if rest.startswith("("):
                depth, i = 0, 0
                for i, ch in enumerate(rest):
                    depth += ch == "("
                    depth -= ch == ")"
                    if depth == 0:
                        break
                n = _call_arity(rest[1:i])
            else:
                n = 0
            assert n == len(cr[name][1]), (os.path.basename(path), name, n, len(cr[name][1]))
            calls += 1
    assert calls >= 25


def test_crystal_sources_are_balanced():
    """A cheap structural check in lieu of a compiler: every block opener has its `end`."""
    opener = re.compile(r"^\s*(?:private |protected |abstract )?(?:def|class|module|struct|lib|enum|macro|if|unless|while|case|begin)\b"
                        r"|\bdo\b(?:\s*\|[^|]*\|)?\s*$|^\s*\{%\s*(?:for|if|begin)\b")
    for path in glob.glob(os.path.join(ROOT, "crystal", "**", "*.cr"), recursive=True):
        depth = 0
        for raw in open(path):
            line = re.sub(r'"(?:[^"\\]|\\.)*"', '""', raw)
            line = re.sub(r"#(?!\{).*", "", line).rstrip()
            if not line.strip():
                continue
            s = line.strip()
            if re.match(r"^(?:private |protected )?abstract def\b", s):
                continue
            if re.match(r"^\{%\s*(?:end)\s*%\}$", s) or s in ("end", "{% end %}") or re.match(r"^end\b", s):
                depth -= 1
                continue
            if re.match(r"^\{%\s*(?:else|elsif)\b", s) or re.match(r"^(?:else|elsif|when|rescue|ensure)\b", s):
                continue
            one_liner = re.search(r"\bend\s*$", s) is not None and not s.startswith("end")
            if opener.search(line) and not one_liner:
                # suffix `if` / `unless` modifiers do not open a block
                if re.match(r"^\s*(?:if|unless|while|case|begin)\b", line) or not re.search(r"\S\s+(?:if|unless)\b", line) \
                        or re.match(r"^\s*(?:private |protected )?(?:def|class|module|struct|lib|enum|macro)\b", line) \
                        or re.search(r"\bdo\b(?:\s*\|[^|]*\|)?\s*$", line) or s.startswith("{%"):
                    depth += 1
            assert depth >= 0, (path, raw)
        assert depth == 0, (path, depth)
