#!/usr/bin/env python
"""glsl2cpp.py -- TEST INFRASTRUCTURE.  Source-to-source recipe that lets g++ compile the reference's
ray-tracing shaders *from where they lie* (/root/reference/resources/shaders) as C++:

    python oracle/ref_shim/glsl2cpp.py /root/reference/resources/shaders  out/stages.inc

Nothing of the reference is copied into this repository: the generated file is a build intermediate
written next to the output library under oracle/_ref/ (git-ignored) and deleted after the compile.

Each shader stage (PathTrace.rgen / .rchit / .rahit / .rmiss, PathTraceShadow.rmiss) becomes one
C++ struct deriving from glsl::StageBase (oracle/ref_shim/glsl_shim.hpp); its global variables become
members, its functions member functions, main() stays main().  The rewriting is purely lexical:

  * `#include "x"` is inlined (once per stage); `#version` / `#extension` lines are dropped
  * unsuffixed floating literals get an `f` (GLSL literals are 32-bit floats)
  * `in` / `out` / `inout` parameter qualifiers become by-value / reference parameters
  * multi-component swizzles `.xyz .xy .rgb` become member calls `.xyz()`
  * vector constructor calls `vec3( a, b, c )` become brace initialisation `vec3{ a, b, c }`, which
    C++ evaluates left to right like GLSL does (the arguments draw random numbers)
  * `x = f(..) + g(..) + h(..);` whose operands are calls is sequenced left to right through
    temporaries (C++ leaves operand order unspecified; the light loops all advance ray.seed)
  * `layout(...)` declarations are bound to the descriptor table of the pipeline:
      rayPayloadEXT T n          -> member `T n;` registered under its location
      rayPayloadInEXT T n        -> `T& n` bound to the caller's payload
      hitAttributeEXT T n        -> member
      uniform / buffer blocks    -> struct + reference to the memory bound at (set, binding); unsized
                                    arrays become pointers; `bool` inside blocks is the 4-byte GLSL bool
      push_constant block        -> struct + one reference member per field
      opaque uniforms            -> reference (or pointer for arrays) to the bound object
      constant_id constants      -> static constexpr
  * `ignoreIntersectionEXT;` sets the ignore flag and returns
"""
import os
import re
import sys

STAGES = [
    ("rgen", "PathTrace.rgen"),
    ("rchit", "PathTrace.rchit"),
    ("rahit", "PathTrace.rahit"),
    ("rmiss", "PathTrace.rmiss"),
    ("rmiss_shadow", "PathTraceShadow.rmiss"),
]

VEC = r"(?:[iu]?vec[234])"


def strip_comments(s):
    s = re.sub(r"/\*.*?\*/", " ", s, flags=re.S)
    return re.sub(r"//[^\n]*", "", s)


def inline_includes(path, seen):
    out = []
    base = os.path.dirname(path)
    for line in strip_comments(open(path).read()).split("\n"):
        m = re.match(r'\s*#include\s+"([^"]+)"', line)
        if m:
            inc = os.path.normpath(os.path.join(base, m.group(1)))
            if inc not in seen:
                seen.add(inc)
                out.append(inline_includes(inc, seen))
            continue
        if re.match(r"\s*#(version|extension)\b", line):
            continue
        out.append(line)
    return "\n".join(out)


def suffix_floats(s):
    pat = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])")
    return pat.sub(lambda m: m.group(1) + "f", s)


def match_paren(s, i):
    """index of the parenthesis closing the one at s[i]"""
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced parentheses")


def brace_constructors(s):
    pat = re.compile(r"\b" + VEC + r"\s*\(")
    pos = 0
    s = list(s)
    while True:
        m = pat.search("".join(s), pos)
        if not m:
            break
        o = m.end() - 1
        c = match_paren("".join(s), o)
        s[o], s[c] = "{", "}"
        pos = o + 1
    return "".join(s)


def split_top(expr, sep):
    parts, depth, cur = [], 0, ""
    for ch in expr:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return parts


def sequence_call_sums(s):
    """`lhs = f(..) + g(..) [+ h(..)];`  ->  temporaries evaluated left to right."""
    out, n = [], [0]

    def rewrite(m):
        lhs, rhs = m.group(1), m.group(2)
        parts = [p.strip() for p in split_top(rhs, "+")]
        if len(parts) < 2 or not all(re.fullmatch(r"[A-Za-z_]\w*\s*\(.*\)", p, flags=re.S) for p in parts):
            return m.group(0)
        if any(re.match(VEC + r"\b", p) for p in parts):
            return m.group(0)
        n[0] += 1
        names = [f"seq{n[0]}_{k}" for k in range(len(parts))]
        decl = " ".join(f"auto {nm} = {p};" for nm, p in zip(names, parts))
        return "{ " + decl + f" {lhs} = " + " + ".join(names) + "; }"

    return re.sub(r"([A-Za-z_][\w.]*)\s*=\s*([^;=]+?\)\s*\+[^;=]+?\));", rewrite, s)


def block_members(body):
    """members of a uniform / buffer block: unsized arrays -> pointers, bool -> 4-byte bool"""
    out, fields = [], []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        m = re.fullmatch(r"(\w+)\s+(\w+)\s*(\[\s*(\w*)\s*\])?", decl)
        if not m:
            raise ValueError("cannot parse block member: " + decl)
        ty, name, arr, dim = m.group(1), m.group(2), m.group(3), m.group(4)
        if ty == "bool":
            ty = "bool32"
        if arr and not dim:
            out.append(f"const {ty}* {name};")
        elif arr:
            out.append(f"{ty} {name}[{dim}];")
        else:
            out.append(f"{ty} {name};")
        fields.append((ty, name))
    return "\n  ".join(out), fields


def bind_layouts(s):
    """rewrites every `layout(...) ... ;` declaration (see the module docstring)"""
    payloads = []

    def declaration(quals, text):
        """text = everything between the layout(...) and the terminating semicolon"""
        sb = (quals.get("set", "0"), quals.get("binding", "0"))
        if "constant_id" in quals:  # layout(constant_id = k) const uint N = v
            return "static constexpr " + re.sub(r"\bconst\b", "", text).strip() + ";"
        mb = re.search(r"\{(.*)\}", text, flags=re.S)
        head = text[:mb.start()] if mb else text
        words = [w for w in head.split() if w not in ("readonly", "writeonly")]
        if words[0] == "rayPayloadEXT":
            payloads.append((quals["location"], words[2]))
            return f"{words[1]} {words[2]}{{}};"
        if words[0] == "rayPayloadInEXT":
            return f"{words[1]}& {words[2]} = *static_cast<{words[1]}*>(payloadIn);"
        assert words[0] in ("uniform", "buffer"), text
        bound = f"bindingPtr({sb[0]}, {sb[1]})"
        if not mb:  # opaque uniform: acceleration structure, image, sampler
            ty, name = words[1], words[2]
            if name.endswith("[]"):
                return f"const {ty}* {name[:-2]} = static_cast<const {ty}*>({bound});"
            return f"const {ty}& {name} = *static_cast<const {ty}*>({bound});"
        block, inst = words[1], text[mb.end():].strip()
        members, fields = block_members(mb.group(1))
        struct = f"struct {block} {{\n  {members}\n}};\n"
        if "push_constant" in quals:
            return struct + "\n".join(
                f"const {ty}& {nm} = static_cast<const {block}*>(pushConstantPtr())->{nm};" for ty, nm in fields)
        if inst.endswith("[]"):
            return struct + f"const {block}* {inst[:-2]} = static_cast<const {block}*>({bound});"
        return struct + f"const {block}& {inst} = *static_cast<const {block}*>({bound});"

    out, pos = [], 0
    for m in re.finditer(r"\blayout\s*\(", s):
        if m.start() < pos:
            continue
        close = match_paren(s, m.end() - 1)
        quals = {}
        for kv in s[m.end():close].split(","):
            k, _, v = kv.partition("=")
            quals[k.strip()] = v.strip() or True
        depth, end = 0, close + 1
        while not (s[end] == ";" and depth == 0):
            depth += {"{": 1, "}": -1}.get(s[end], 0)
            end += 1
        out.append(s[pos:m.start()])
        out.append(declaration(quals, s[close + 1:end].strip()))
        pos = end + 1
    out.append(s[pos:])
    s = "".join(out)
    s = re.sub(r"\bhitAttributeEXT\s+(\w+)\s+(\w+)\s*;", r"\1 \2{};", s)
    cases = " ".join(f"case {loc}: return &{name};" for loc, name in payloads)
    s += f"\nvoid* payloadAt(int location) override {{ switch (location) {{ {cases} default: return nullptr; }} }}\n"
    return s


def parameters(s):
    s = re.sub(r"\b(?:inout|out)\s+(\w+)\s+(\w+)", r"\1& \2", s)
    return re.sub(r"\bin\s+(\w+)\s+(\w+)", r"\1 \2", s)


def translate(path):
    s = inline_includes(path, {os.path.normpath(path)})
    s = suffix_floats(s)
    s = bind_layouts(s)
    s = parameters(s)
    s = re.sub(r"\.(xyz|xy|rgb)\b(?!\s*\()", r".\1()", s)
    s = brace_constructors(s)
    s = sequence_call_sums(s)
    s = re.sub(r"\bignoreIntersectionEXT\s*;", "{ ignoreIntersection_ = true; return; }", s)
    return re.sub(r"\n\s*\n+", "\n", s)


def main():
    src, out = sys.argv[1], sys.argv[2]
    parts = ["// GENERATED by oracle/ref_shim/glsl2cpp.py from " + src + " -- build intermediate, do not commit\n"]
    for name, fn in STAGES:
        body = translate(os.path.join(src, fn))
        parts.append(f"struct Stage_{name} : StageBase {{\n  using StageBase::StageBase;\n{body}\n}};\n#undef M_PI\n")
    with open(out, "w") as f:
        f.write("\n".join(parts))


if __name__ == "__main__":
    main()
