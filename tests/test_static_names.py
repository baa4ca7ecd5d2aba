"""CPU: no function of the package, the bench, the entry points or the tests reads a name that nothing defines.
Most of the product only executes on the GPU box; this catches the typo class that would otherwise surface there
(the image has no pyflakes; tools/undefined_names.py is a symtable walk)."""
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import undefined_names  # noqa: E402


def test_no_undefined_names():
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    for sub in ("prodsearch_b200", "oracle", "tests", "profiles", "tests/golden"):
        files += sorted(glob.glob(os.path.join(ROOT, sub, "*.py")))
    bad = [(os.path.relpath(f, ROOT),) + b for f in files for b in undefined_names.check(f)]
    assert not bad, bad


def test_checker_sees_an_undefined_name(tmp_path):
    p = tmp_path / "m.py"
    p.write_text("import os\n\ndef f():\n    return os.sep + missing_name\n")
    assert undefined_names.check(str(p)) == [("f", "missing_name")]
