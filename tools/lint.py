"""Undefined-name check of every Python source (the GPU paths cannot run in the CPU-only build container, so a
NameError would otherwise surface only on the GPU box): compiles each file and reports names that are loaded but never
bound in the module / builtins (a tiny pyflakes substitute; pyflakes is not in the image)."""
import ast
import builtins
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class V(ast.NodeVisitor):
    def __init__(self):
        self.scopes = [set(dir(builtins)) | {'__file__', '__name__', '__doc__'}]
        self.problems = []

    def bind_targets(self, node, scope):
        for n in ast.walk(node):
            if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
                scope.add(n.id)
            elif isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
                scope.add(n.name)
            elif isinstance(n, (ast.Import, ast.ImportFrom)):
                for a in n.names:
                    scope.add((a.asname or a.name).split('.')[0])
            elif isinstance(n, ast.ExceptHandler) and n.name:
                scope.add(n.name)
            elif isinstance(n, (ast.Global, ast.Nonlocal)):
                scope.update(n.names)
            elif isinstance(n, ast.arg):
                scope.add(n.arg)

    def visit_scope(self, node, body):
        scope = set()
        if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
            a = node.args
            for x in a.posonlyargs + a.args + a.kwonlyargs + ([a.vararg] if a.vararg else []) + ([a.kwarg] if a.kwarg else []):
                scope.add(x.arg)
        for st in body if isinstance(body, list) else [body]:
            self.bind_targets(st, scope)
        self.scopes.append(scope)
        for st in body if isinstance(body, list) else [body]:
            self.visit(st)
        self.scopes.pop()

    def visit_Module(self, node):
        self.visit_scope(node, node.body)

    def visit_FunctionDef(self, node):
        for d in node.decorator_list + node.args.defaults + [x for x in node.args.kw_defaults if x]:
            self.visit(d)
        self.visit_scope(node, node.body)
    visit_AsyncFunctionDef = visit_FunctionDef

    def visit_Lambda(self, node):
        self.visit_scope(node, node.body)

    def visit_ClassDef(self, node):
        for d in node.decorator_list + node.bases:
            self.visit(d)
        self.visit_scope(node, node.body)

    def visit_Name(self, node):
        if isinstance(node.ctx, ast.Load) and not any(node.id in s for s in self.scopes):
            self.problems.append((node.lineno, node.id))


def main(paths):
    bad = 0
    for p in paths:
        src = open(p).read()
        tree = ast.parse(src, p)
        v = V()
        v.visit(tree)
        for line, name in v.problems:
            print('%s:%d: undefined name %r' % (os.path.relpath(p, ROOT), line, name))
            bad += 1
    return bad


if __name__ == '__main__':
    files = sys.argv[1:]
    if not files:
        for d in ('ess_b200', 'tests', 'tools', 'oracle', '.'):
            dd = os.path.join(ROOT, d)
            files += [os.path.join(dd, f) for f in sorted(os.listdir(dd)) if f.endswith('.py')]
    sys.exit(1 if main(files) else 0)
