"""The documents must not name things that do not exist: every `tritd_*` entry point and every `tritd.<name>` binding that
INTEGRATION.md, DESIGN.md or README.md mention is declared in include/tritd.h / defined in the Python mirror, every
`tests/...`, `tools/...`, `profiles/...`, `oracle/...` path they cite is in the tree, and every TRITD_* environment switch
they document is read somewhere."""
import glob
import os
import re

import tritd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOCS = ["INTEGRATION.md", "DESIGN.md", "README.md", os.path.join("profiles", "README.md")]
HEADER = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "tritd.h")).read(), flags=re.S)


def _text(doc):
    return open(os.path.join(ROOT, doc)).read()


def test_named_entry_points_exist():
    declared = set(re.findall(r"\b(tritd_[a-z0-9_A-Z]+)\s*\(", HEADER)) | set(re.findall(r"\b(tritd_[a-z_]+)\b(?=;|\s*\{|\s+\w+;)", HEADER))
    types = set(re.findall(r"\}\s*(tritd_[a-z_]+);", HEADER)) | set(re.findall(r"typedef struct (tritd_[a-z_]+)", HEADER)) | {"tritd_print_fn"}
    internal = {"tritd_oracle", "tritd_oracle_mt", "tritd_oracle_sharded", "tritd_b200", "tritd_problem_create_level"}
    for doc in DOCS:
        for name in set(re.findall(r"`(tritd_[a-z0-9_]+)", _text(doc))):
            if name in internal or name.startswith("tritd_oracle"):
                continue
            candidates = {name} | {name + suf for suf in ("_f64", "_dev_f64")} | {n for n in declared if n.startswith(name)}
            assert candidates & (declared | types), (doc, name)


def test_named_python_bindings_exist():
    for doc in DOCS:
        for name in set(re.findall(r"`tritd\.([A-Za-z_][A-Za-z0-9_]*(?:\.[A-Za-z_][A-Za-z0-9_]*)?)", _text(doc))):
            if name == "h":                       # the header include/tritd.h
                continue
            obj = tritd
            for part in name.split("."):
                assert hasattr(obj, part), (doc, name)
                obj = getattr(obj, part)


def test_cited_paths_exist():
    for doc in DOCS:
        base = os.path.join(ROOT, "profiles") if doc.startswith("profiles") else ROOT
        for path in set(re.findall(r"`((?:tests|tools|profiles|oracle|include)/[A-Za-z0-9_./{},*-]+)`", _text(doc))):
            if path.endswith("/"):
                assert os.path.isdir(os.path.join(ROOT, path)), (doc, path)
                continue
            if "matlab_" in path or "dump_in_" in path or "_ref" in path:      # produced on demand (MATLAB dump, reference install)
                continue
            path = path.split("::")[0]
            pats = [path]
            m = re.search(r"\{([^}]*)\}", path)
            if m:
                pats = [path[:m.start()] + alt + path[m.end():] for alt in m.group(1).split(",")]
            for pat in pats:
                assert glob.glob(os.path.join(ROOT, pat)), (doc, pat)
        if doc.startswith("profiles"):
            for f in set(re.findall(r"`(r0[12][a-z]?_[A-Za-z0-9_.{},*-]+|fused_traffic\.json)`", _text(doc))):
                pats = [f]
                m = re.search(r"\{([^}]*)\}", f)
                if m:
                    pats = [f[:m.start()] + alt + f[m.end():] for alt in m.group(1).split(",")]
                for pat in pats:
                    assert glob.glob(os.path.join(base, pat)), (doc, pat)


def test_documented_environment_switches_are_read():
    src = ""
    for pat in ("triple-tensor-decomposition-with-admm_b200/csrc/*", "triple-tensor-decomposition-with-admm_b200/tritd/*.py", "bench.py", "tools/*.py"):
        for f in glob.glob(os.path.join(ROOT, pat)):
            if os.path.isfile(f) and not f.endswith(".so"):
                src += open(f, errors="ignore").read()
    for name in set(re.findall(r"\| `(TRITD_[A-Z0-9_]+)", _text("INTEGRATION.md"))):
        assert name in src, name
