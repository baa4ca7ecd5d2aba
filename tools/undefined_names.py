"""Poor man's pyflakes (none is installed in the image): report names a function reads from the module / builtin
scope that neither defines -- the typo class a GPU-only code path would otherwise hide until it runs on the box."""
import builtins
import symtable
import sys


def check(path):
    src = open(path).read()
    top = symtable.symtable(src, path, "exec")
    module_names = set(top.get_identifiers())
    bad = []

    def walk(t):
        for s in t.get_symbols():
            if s.is_referenced() and (s.is_global() or (t is top and not s.is_assigned() and not s.is_imported()
                                                         and not s.is_namespace())):
                n = s.get_name()
                if n not in module_names and not hasattr(builtins, n) and n not in ("__file__", "__name__"):
                    bad.append((t.get_name(), n))
                elif t is not top and n in module_names:
                    m = top.lookup(n)
                    if not (m.is_assigned() or m.is_imported() or m.is_namespace()) and not hasattr(builtins, n):
                        bad.append((t.get_name(), n))
        for c in t.get_children():
            walk(c)
    walk(top)
    return bad


if __name__ == "__main__":
    rc = 0
    for p in sys.argv[1:]:
        for scope, name in check(p):
            print("%s: undefined name %r in %s" % (p, name, scope))
            rc = 1
    sys.exit(rc)
