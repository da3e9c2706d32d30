"""Extracts, with `ast` (the config modules cannot be imported: datasets, pycocotools, ...), the
keyword dictionaries of every `losses.__dict__[<name>](**{...})` / `decode.__dict__[<name>](**{...})`
call for the classes this repo provides, from the reference's own config files:

    3.detection_training/**/{train,test}_config.py      RetinaLoss, FCOSLoss, RetinaDecoder,
                                                        FCOSDecoder, DETRDecoder, DINODETRDecoder
    10.face_detection_training/**/{train,test}_config.py  RetinaFaceLoss, RetinaFaceDecoder

and writes tests/golden/config_kwargs.json: one entry per (file, line, class) with the evaluated
kwargs.  Run in the development container only:  python tests/golden/make_config_kwargs.py
"""
import ast
import glob
import json
import os

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
CLASSES = {'RetinaLoss', 'FCOSLoss', 'RetinaDecoder', 'FCOSDecoder', 'DETRDecoder',
           'DINODETRDecoder', 'RetinaFaceLoss', 'RetinaFaceDecoder'}


def literal(node, env):
    """Evaluates the small expression language the configs use inside the dicts: literals,
    arithmetic on literals (2**(1.0 / 3.0)), and names bound to literals earlier in the file."""
    return eval(compile(ast.Expression(node), '<cfg>', 'eval'), {'__builtins__': {}}, dict(env))


def module_constants(tree):
    env = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and \
                isinstance(node.targets[0], ast.Name):
            try:
                env[node.targets[0].id] = literal(node.value, env)
            except Exception:
                pass
    return env


def calls_in(path):
    src = open(path).read()
    tree = ast.parse(src)
    env = module_constants(tree)
    out = []
    for node in ast.walk(tree):
        if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Subscript)):
            continue
        sub = node.func
        if not (isinstance(sub.value, ast.Attribute) and sub.value.attr == '__dict__' and
                isinstance(sub.value.value, ast.Name) and sub.value.value.id in ('losses', 'decode')):
            continue
        key = sub.slice
        if not (isinstance(key, ast.Constant) and key.value in CLASSES):
            continue
        kwargs = {}
        for kw in node.keywords:
            if kw.arg is None:                       # **{...}
                kwargs.update(literal(kw.value, env))
            else:
                kwargs[kw.arg] = literal(kw.value, env)
        out.append({'file': os.path.relpath(path, REF), 'line': node.lineno,
                    'module': sub.value.value.id, 'class': key.value, 'kwargs': kwargs})
    return out


def main():
    entries = []
    for pattern in ('3.detection_training/**/*_config.py', '10.face_detection_training/**/*_config.py'):
        for path in sorted(glob.glob(os.path.join(REF, pattern), recursive=True)):
            entries.extend(calls_in(path))
    with open(os.path.join(HERE, 'config_kwargs.json'), 'w') as f:
        json.dump(entries, f, indent=1, sort_keys=True)
    by = {}
    for e in entries:
        by[e['class']] = by.get(e['class'], 0) + 1
    print(len(entries), 'calls:', by)


if __name__ == '__main__':
    main()
