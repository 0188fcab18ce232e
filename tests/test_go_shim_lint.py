"""Static checks of the Go sources (go/): there is no Go toolchain in this image, so the shim cannot be compiled here.
What CAN be checked without one, and is: the files tokenise as Go, brackets balance, no import is unused (a compile
error in Go), and every C.tfhe_* call in the cgo shim names a function that include/tfhe_b200.h declares, with the
declared number of arguments, and every C type it names exists in the header."""
import os
import re

import pytest

pygments = pytest.importorskip("pygments")
from pygments.lexers import GoLexer  # noqa: E402
from pygments.token import Comment, Error, String  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO_FILES = [os.path.join(ROOT, "go", "tfheb200", "tfheb200.go"), os.path.join(ROOT, "go", "tfheb200", "wire.go"),
            os.path.join(ROOT, "go", "cmd", "mkgolden", "main.go")]


def _code_tokens(path):
    src = open(path).read()
    toks = list(GoLexer().get_tokens(src))
    return src, [(t, v) for t, v in toks if t not in Comment and not (t in Comment.Multiline)]


@pytest.mark.parametrize("path", GO_FILES, ids=[os.path.basename(p) for p in GO_FILES])
def test_go_file_tokenises_and_brackets_balance(path):
    src, toks = _code_tokens(path)
    assert not [v for t, v in toks if t in Error], "lexer errors"
    stack, pairs = [], {")": "(", "]": "[", "}": "{"}
    for t, v in toks:
        if t in String or t in Comment:
            continue
        for ch in v:
            if ch in "([{":
                stack.append(ch)
            elif ch in ")]}":
                assert stack and stack.pop() == pairs[ch], "unbalanced %r in %s" % (ch, path)
    assert not stack
    assert re.search(r"^package \w+", src, re.M)


def _strip_comments(src):
    src = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


@pytest.mark.parametrize("path", GO_FILES, ids=[os.path.basename(p) for p in GO_FILES])
def test_go_imports_are_all_used(path):
    code = _strip_comments(open(path).read())
    m = re.search(r"\nimport \(\n(.*?)\n\)", code, re.S)
    assert m, "no import block"
    body = code[m.end():]
    for line in m.group(1).splitlines():
        line = line.strip()
        if not line:
            continue
        mm = re.match(r'(?:(\w+)\s+)?"([^"]+)"', line)
        assert mm, line
        name = mm.group(1) or mm.group(2).split("/")[-1]
        assert re.search(r"\b%s\." % re.escape(name), body), "import %s is never used in %s (a Go compile error)" % (mm.group(2), path)


def _header_functions():
    hdr = _strip_comments(open(os.path.join(ROOT, "include", "tfhe_b200.h")).read())
    funcs = {}
    for m in re.finditer(r"\b(tfhe_\w+)\s*\(([^;{}]*?)\)\s*;", hdr, re.S):
        params = m.group(2).strip()
        funcs[m.group(1)] = 0 if params in ("", "void") else params.count(",") + 1
    types = set(re.findall(r"}\s*(tfhe_\w+)\s*;", hdr)) | set(re.findall(r"typedef\s+struct\s+\w+\s+(tfhe_\w+)\s*;", hdr))
    return funcs, types


def _call_args(code, start):
    """Number of top-level arguments of the call whose '(' is at code[start]."""
    depth, n, any_tok = 0, 0, False
    for i in range(start, len(code)):
        ch = code[i]
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
            if depth == 0:
                return n + 1 if any_tok else 0
        elif ch == "," and depth == 1:
            n += 1
        elif depth >= 1 and not ch.isspace():
            any_tok = True
    raise AssertionError("unterminated call")


def test_cgo_calls_match_the_header():
    funcs, types = _header_functions()
    assert "tfhe_gate_batch" in funcs and funcs["tfhe_gate_batch"] == 8 and "tfhe_params" in types and "tfhe_ctx" in types
    code = _strip_comments(open(GO_FILES[0]).read())
    code = code[code.index('import "C"'):]          # after the cgo preamble
    calls = list(re.finditer(r"\bC\.(tfhe_\w+)\s*\(", code))
    assert len(calls) >= 15
    for m in calls:
        name = m.group(1)
        if name in types:                            # a conversion such as C.tfhe_params{...} never has '(' — but be safe
            continue
        assert name in funcs, "C.%s is not declared in include/tfhe_b200.h" % name
        got = _call_args(code, m.end() - 1)
        assert got == funcs[name], "C.%s called with %d arguments, the header declares %d" % (name, got, funcs[name])
    for name in set(re.findall(r"\bC\.(tfhe_\w+)\b(?!\s*\()", code)):
        assert name in types or name in funcs, "C.%s is not a type or function of the header" % name


def test_gate_opcodes_match_the_header():
    """The Go Op constants (iota order) must be the header's tfhe_op values."""
    hdr = open(os.path.join(ROOT, "include", "tfhe_b200.h")).read()
    enum = re.search(r"typedef enum\s*{(.*?)}\s*tfhe_op\s*;", hdr, re.S)
    assert enum
    c_ops = [(m.group(1), int(m.group(2))) for m in re.finditer(r"TFHE_OP_(\w+)\s*=\s*(\d+)", _strip_comments(enum.group(1)))]
    go = _strip_comments(open(GO_FILES[0]).read())
    block = re.search(r"const \(\n\s*NAND Op = iota\n(.*?)\n\)", go, re.S)
    assert block
    go_ops = ["NAND"] + [ln.strip() for ln in block.group(1).splitlines() if ln.strip()]
    assert [n for n, _ in sorted(c_ops, key=lambda kv: kv[1])] == go_ops


REFERENCE = "/root/reference"
REF_PKGS = ("cloudkey", "params", "tlwe", "trgsw", "trlwe", "key", "evaluator", "gates", "lut", "poly")


def _reference_symbols(pkg):
    """{name: number of parameters or None} for the package-level funcs, types, consts and vars of a reference package."""
    syms = {}
    d = os.path.join(REFERENCE, pkg)
    for fn in sorted(os.listdir(d)):
        if not fn.endswith(".go") or fn.endswith("_test.go"):
            continue
        code = _strip_comments(open(os.path.join(d, fn)).read())
        for m in re.finditer(r"^func (\w+)\s*\(", code, re.M):
            syms[m.group(1)] = _call_args(code, m.end() - 1)
        for m in re.finditer(r"^type (\w+)\b", code, re.M):
            syms.setdefault(m.group(1), None)
        for m in re.finditer(r"^\s*(?:var|const)\s+(\w+)\b", code, re.M):
            syms.setdefault(m.group(1), None)
        for blk in re.finditer(r"^(?:const|var) \((.*?)^\)", code, re.M | re.S):
            for m in re.finditer(r"^\s*(\w+)\b", blk.group(1), re.M):
                syms.setdefault(m.group(1), None)
    return syms


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present (it never is on the GPU box)")
@pytest.mark.parametrize("path", GO_FILES, ids=[os.path.basename(p) for p in GO_FILES])
def test_reference_symbols_used_by_the_go_sources_exist(path):
    """Every package-level identifier of the reference that the Go sources name (cloudkey.NewCloudKey, params.Torus,
    gates.MUX, ...) must exist in the reference checkout, and calls must pass as many arguments as the function declares."""
    code = _strip_comments(open(path).read())
    code = re.sub(r'"(?:[^"\\]|\\.)*"', '""', code)                 # drop string literals
    imports = re.search(r"\nimport \(\n(.*?)\n\)", open(path).read(), re.S).group(1)
    used = [p for p in REF_PKGS if re.search(r'go-tfhe/%s"' % p, imports)]
    body = code[code.index("\nimport ("):]
    body = body[body.index("\n)") + 2:]
    checked = 0
    for pkg in used:
        syms = _reference_symbols(pkg)
        for m in re.finditer(r"(?<![\w.])%s\.(\w+)" % pkg, body):
            name = m.group(1)
            assert name in syms, "%s.%s does not exist in the reference (%s)" % (pkg, name, path)
            checked += 1
            after = body[m.end():m.end() + 1]
            if after == "(" and syms[name] is not None:
                got = _call_args(body, m.end())
                assert got == syms[name], "%s.%s called with %d arguments, the reference declares %d" % (pkg, name, got, syms[name])
    assert checked > 0 or not used


def _struct_fields_and_methods(code, any_case=False):
    names = set(re.findall(r"^func \([^)]*\) (\w+)\(", code, re.M))
    first = r"[A-Za-z]" if any_case else r"[A-Z]"
    for blk in re.finditer(r"^type \w+ struct \{(.*?)^\}", code, re.M | re.S):
        for ln in blk.group(1).splitlines():
            m = re.match(r"\s*(%s\w*(?:\s*,\s*%s\w*)*)\s+\S" % (first, first), ln)
            if m:
                names |= set(x.strip() for x in m.group(1).split(","))
    return names


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present (it never is on the GPU box)")
def test_fields_and_methods_named_by_the_go_sources_exist():
    """Every exported selector on a value (ck.BootstrappingKey, row.TRLWEFFT, t.A.Coeffs, ev.BootstrapLUTAssign, ...) must be
    a struct field or method of the reference, or one defined in go/ itself; a handful of standard-library methods aside."""
    known = set()
    for root, _, files in os.walk(REFERENCE):
        for fn in files:
            if fn.endswith(".go"):
                known |= _struct_fields_and_methods(_strip_comments(open(os.path.join(root, fn)).read()))
    for path in GO_FILES:
        known |= _struct_fields_and_methods(_strip_comments(open(path).read()), any_case=True)
    stdlib_methods = {"Lock", "Unlock", "Bytes", "Write", "WriteString", "Uint32", "Uint64"}
    packages = set(REF_PKGS) | {"C", "fmt", "runtime", "sync", "unsafe", "bytes", "binary", "errors", "crc32", "io", "math", "os", "flag",
                                "rand", "filepath", "strings", "tfheb200"}
    for path in GO_FILES:
        code = re.sub(r'"(?:[^"\\]|\\.)*"', '""', _strip_comments(open(path).read()))
        for m in re.finditer(r"(\w+|\)|\])\.([A-Z]\w*)", code):
            if m.group(1) in packages:
                continue
            assert m.group(2) in known or m.group(2) in stdlib_methods, "%s: .%s is not a field or method of the reference or of go/" % (
                os.path.basename(path), m.group(2))


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present (it never is on the GPU box)")
def test_shim_does_not_create_an_import_cycle():
    """gates and evaluator (b200 build) import tfheb200; tfheb200 must therefore not reach them through its own imports."""
    def imports_of(pkg_dir):
        out = set()
        for fn in os.listdir(pkg_dir):
            if fn.endswith(".go") and not fn.endswith("_test.go"):
                out |= set(re.findall(r'"github.com/thedonutfactory/go-tfhe/(\w+)"', open(os.path.join(pkg_dir, fn)).read()))
        return out
    seen, todo = set(), set(re.findall(r'"github.com/thedonutfactory/go-tfhe/(\w+)"', open(GO_FILES[0]).read() + open(GO_FILES[1]).read()))
    while todo:
        p = todo.pop()
        if p in seen:
            continue
        seen.add(p)
        todo |= imports_of(os.path.join(REFERENCE, p)) - seen
    assert "gates" not in seen and "evaluator" not in seen, sorted(seen)
