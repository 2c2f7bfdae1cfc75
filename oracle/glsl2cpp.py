#!/usr/bin/env python
"""ORACLE — TEST INFRASTRUCTURE ONLY. Turns one of the reference's GLSL shaders, READ WHERE IT LIES under /root/reference/Shaders,
into text g++ accepts as the body of a C++ struct (oracle/glsl_compat.h supplies GLSL's types and built-ins). The output goes to
oracle/_ref/gen/ (git-ignored build output, deleted by the Makefile rule once the library is linked): no shader text is committed to or kept in this repository.

The rewrite is mechanical and token-level; it never touches an expression:
  1. `#include <...>` resolved against the Shaders directory (textually, as the reference's shaderc includer does,
     Src/Shader.cpp:40-57); `#version` / `#extension` lines dropped;
  2. the C preprocessor (`g++ -E -P`) expands the shader's own macros (the bindless sugar of Bindless/GlobalHeap.glsl
     included), so that every declaration appears in plain GLSL;
  3. interface declarations lose their `layout(...)` and storage qualifiers: `uniform sampler2D heap[];` -> `sampler2D* heap;`,
     `uniform Block {..} name[];` -> `struct Block {..}; Block* name;`, `in vec3 x;` -> `vec3 x;`; an unsized array member
     `T arr[];` becomes `T* arr;`, a sized array `T a[N];` the value type `arr<T, N> a;`
  4. `out T p` / `inout T p` parameters become `T& p`, `in T p` becomes `T p`;
  5. floating-point literals get the `f` suffix (a GLSL `1.0` is a float, a C++ `1.0` a double);
  6. vector / matrix constructor calls are written with braces, `vec3(a, b, c)` -> `vec3{a, b, c}`: GLSL evaluates arguments left
     to right, C++ only promises that inside a braced list (SSAO.glsl:37 draws `vec3(rng(), rng(), rng())`).

Patches (`--patch 'regex=>replacement'`) apply to the ROOT shader's own text before step 1 (so they can add an include). They exist for the bit-rot SURVEY.md 8(c-bis) lists (a call
with a stale signature, a missing include) and are spelled out in oracle/Makefile, one per defect, citing the reconciliation rule.

  python oracle/glsl2cpp.py <shader relative to Shaders/> <out.inc> [--root DIR] [--patch 'a=>b']... [--prelude FILE] [-DNAME[=V]]...
"""
from __future__ import annotations

import os
import re
import subprocess
import sys


def inline_includes(root: str, rel: str, seen: list, patches=()) -> str:
    path = os.path.normpath(os.path.join(root, rel))
    seen.append(path)
    out = []
    src = open(path, encoding="utf-8", errors="replace").read()
    for p in patches:
        pat, rep = p.split("=>", 1)
        src, n = re.subn(pat, rep.replace("\\n", "\n"), src)
        if n == 0:
            raise SystemExit("glsl2cpp: patch did not apply: " + pat)
    for line in src.splitlines():
        m = re.match(r'\s*#\s*include\s*[<"]([^>"]+)[>"]', line)
        if m and os.path.isfile(os.path.normpath(os.path.join(root, m.group(1)))):
            out.append(inline_includes(root, m.group(1), seen))
            continue  # (an include that does not resolve stays: it sits in a branch the preprocessor never takes for a shader)
        if re.match(r"\s*#\s*(version|extension)\b", line):
            continue
        out.append(line)
    return "\n".join(out) + "\n"


QUAL = r"(?:uniform|buffer|readonly|writeonly|coherent|restrict|smooth|flat|noperspective|in|out)"


def declarations(text: str) -> str:
    # layout(...) [qualifiers] -> marker
    text = re.sub(r"\blayout\s*\((?:[^()]|\([^()]*\))*\)\s*(?:%s\s+)*(?:%s\b)?" % (QUAL, QUAL), "GLSL_DECL ", text)
    # qualifiers without a layout at the start of a global declaration
    text = re.sub(r"(?m)^(\s*)(?:%s\s+)+(?=\w+\s+\w+\s*(?:\[\s*\])?\s*;)" % QUAL, r"\1GLSL_DECL ", text)

    def block(m):
        name, body, inst, arr = m.group(1), m.group(2), m.group(3), m.group(4)
        body = re.sub(r"(\w+)\s+(\w+)\s*\[\s*\]\s*;", r"\1* \2;", body)  # unsized array member
        return "struct %s {%s}; %s%s %s;" % (name, body, name, "*" if arr else "", inst)

    text = re.sub(r"GLSL_DECL\s+(\w+)\s*\{([^{}]*)\}\s*(\w+)\s*(\[\s*\])?\s*;", block, text)
    text = re.sub(r"GLSL_DECL\s+(\w+)\s+(\w+)\s*\[\s*(\d+)\s*\]\s*;", r"\1 \2[\3];", text)  # sized interface array: arrays() below
    text = re.sub(r"GLSL_DECL\s+(\w+)\s+(\w+)\s*\[\s*\]\s*;", r"\1* \2;", text)
    text = re.sub(r"GLSL_DECL\s+(\w+)\s+(\w+)\s*;", r"\1 \2;", text)
    text = re.sub(r"GLSL_DECL\s*;", "", text)  # layout(local_size_x = ..) in;
    if "GLSL_DECL" in text:
        bad = text[text.index("GLSL_DECL"):][:160]
        raise SystemExit("glsl2cpp: a declaration form this translator does not know: " + bad)
    return text


def arrays(text: str) -> str:
    """`T name[N];` -> `arr<T, N> name;`: a GLSL array is a value (assignable, copied with the struct that holds it)."""
    return re.sub(r"\b(u?vec[234]|ivec2|mat[34]|float|int|uint)\s+(\w+)\s*\[\s*(\d+)\s*\]\s*;", r"arr<\1, \3> \2;", text)


def parameters(text: str) -> str:
    text = re.sub(r"\b(?:inout|out)\s+(\w+)\s+(\w+)", r"\1& \2", text)
    text = re.sub(r"\bin\s+(\w+)\s+(\w+)", r"\1 \2", text)
    return text


CTOR = re.compile(r"\b(u?vec[234]|ivec2|mat[34])\s*\(")


def constructors(text: str) -> str:
    """T(a, b, ..) -> T{a, b, ..} for the vector / matrix types: left-to-right evaluation of the arguments, as GLSL specifies."""
    out, i = [], 0
    while True:
        m = CTOR.search(text, i)
        if not m:
            out.append(text[i:])
            return "".join(out)
        # a declaration `vec3 name(` never matches (the type is followed by a name); find the matching parenthesis
        depth, j = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(text[j], 0)
            j += 1
        out.append(text[i:m.start()] + m.group(1) + "{" + constructors(text[m.end():j - 1]) + "}")
        i = j


def literals(text: str) -> str:
    num = r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])"
    return re.sub(num, r"\1f", text)


def main(argv):
    root = "/root/reference/Shaders"
    patches, patches_all, defines, prelude = [], [], [], ""
    pos = []
    it = iter(argv)
    for a in it:
        if a == "--root": root = next(it)
        elif a == "--patch": patches.append(next(it))
        elif a == "--patch-all": patches_all.append(next(it))  # applied after include resolution (the expression lives in an included file)
        elif a == "--prelude": prelude += open(next(it)).read() + "\n"
        elif a.startswith("-D"): defines.append(a)
        else: pos.append(a)
    rel, out = pos
    seen = []
    # the prelude goes through the same include resolution (it may name headers the bit-rotted shader forgot)
    text = ""
    for line in prelude.splitlines():
        m = re.match(r'\s*#\s*include\s*[<"]([^>"]+)[>"]', line)
        text += inline_includes(root, m.group(1), seen) if m else line + "\n"
    text += inline_includes(root, rel, seen, patches)
    for p in patches_all:
        pat, rep = p.split("=>", 1)
        text, n = re.subn(pat, rep, text)
        if n == 0:
            raise SystemExit("glsl2cpp: patch did not apply: " + pat)
    text = re.sub(r"//[^\n]*", "", text)  # comments may hold apostrophes the C preprocessor trips over
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    pre = subprocess.run(["g++", "-E", "-P", "-undef", "-nostdinc", "-x", "c", "-"] + defines, input=text, capture_output=True, text=True)
    if pre.returncode != 0:
        raise SystemExit("glsl2cpp: preprocessor failed:\n" + pre.stderr)
    text = literals(constructors(parameters(arrays(declarations(pre.stdout)))))
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    with open(out, "w") as f:
        f.write("// generated by oracle/glsl2cpp.py from %s (and %d includes) -- build output, do not commit\n" % (rel, len(seen) - 1))
        f.write(text)
    print("glsl2cpp: %s -> %s (%d files)" % (rel, out, len(seen)))


if __name__ == "__main__":
    main(sys.argv[1:])
