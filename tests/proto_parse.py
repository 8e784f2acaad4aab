"""Normalised C prototypes out of a public header (TEST INFRASTRUCTURE): `name -> "ret(type, type, ...)"` with comments,
parameter names and spacing removed, so two headers can be compared declaration by declaration."""
import re

_KEYWORDS = {"const", "unsigned", "signed", "int", "long", "short", "char", "double", "float", "void", "size_t", "struct", "enum"}


def _norm_type(t):
    t = re.sub(r"\s+", " ", t.strip())
    t = re.sub(r"\s*\*\s*", "*", t)
    return t


def _strip_name(param):
    param = param.strip()
    if param in ("void", "..."):
        return param
    m = re.match(r"^(.*?)([A-Za-z_]\w*)\s*(\[\s*\])?$", param, flags=re.S)
    if m and m.group(1).strip() and m.group(2) not in _KEYWORDS and not m.group(2).endswith("_t"):
        base = m.group(1) + ("*" if m.group(3) else "")
        return _norm_type(base)
    return _norm_type(param)


def prototypes(text):
    src = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"^\s*#.*$", " ", src, flags=re.M)
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?[\s\*])((?:fft|fftb200)_\w+)\s*\(([^;{()]*)\)\s*;", src):
        ret, name, params = m.group(1), m.group(2), m.group(3)
        if "typedef" in ret or "static" in ret or "return" in ret:
            continue
        plist = [_strip_name(p) for p in params.split(",")] if params.strip() else ["void"]
        out[name] = f"{_norm_type(ret)}({', '.join(plist)})"
    return out


def constants(text):
    src = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    out = {}
    for m in re.finditer(r"^\s*#\s*define\s+(FFT_\w+)\s+(\(?[-\w<\s|x()]+\)?)\s*$", src, flags=re.M):
        out[m.group(1)] = re.sub(r"\s+", "", m.group(2))
    for m in re.finditer(r"\b(FFT_[A-Z0-9_]+)\s*=\s*([-\w<\s]+?)\s*[,}]", src):
        out[m.group(1)] = re.sub(r"\s+", "", m.group(2))
    return out
