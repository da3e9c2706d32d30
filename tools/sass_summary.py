"""Per-kernel SASS evidence of what the shipped libb200det.so contains (run anywhere nvcc's cuobjdump
is installed; no GPU needed):  python tools/sass_summary.py > profiles/r02_sass_summary.txt

Counts, per sm_100a kernel, the mnemonics that show how it talks to the hardware: packed FP32
(FFMA2 / FMUL2 / FADD2), 128-bit streaming loads (LDG.E...128), bulk TMA copies (UBLKCP) and their
mbarriers (SYNCS), warp votes (VOTE), cluster / distributed-shared-memory traffic (UCGABAR, ST/LD
with .CLUSTER / mapa), programmatic dependent launch (ACQBULK / PREEXIT-style griddepcontrol), and
tensor-core ops (none expected: nothing on this path is a contraction)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'simpleaicv-pytorch-imagenet-coco-training_b200', 'libb200det.so')
PATTERNS = collections.OrderedDict([
    ('FFMA2', r'\bFFMA2\b'), ('FMUL2', r'\bFMUL2\b'), ('FADD2', r'\bFADD2\b'),
    ('LDG.128', r'\bLDG\.[A-Z0-9.]*128'), ('STG.128', r'\bSTG\.[A-Z0-9.]*128'),
    ('UBLKCP', r'\bUBLKCP'), ('SYNCS', r'\bSYNCS'), ('VOTE', r'\bVOTE'),
    ('FMNMX3', r'\bFMNMX3'), ('MUFU', r'\bMUFU'),
    ('CLUSTER', r'UCGABAR|\.CLUSTER|\bMAPA\b|CGA'), ('PDL', r'ACQBULK|PREEXIT|DEPBAR\.LE.*SB'),
    ('ATOMS', r'\bATOMS'), ('tensor', r'\b(HMMA|IMMA|QGMMA|UTCMMA|UTCHMMA|TCGEN)'),
])


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    arch = set(re.findall(r'arch = (sm_\w+)', out))
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None or '/*' not in line:
            continue
        kernels[cur]['instructions'] += 1
        for name, pat in PATTERNS.items():
            if re.search(pat, line):
                kernels[cur][name] += 1
    demangle = subprocess.run(['c++filt'], input='\n'.join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f'# cuobjdump -sass of {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels, architectures {sorted(arch)}')
    cols = ['instructions'] + list(PATTERNS)
    print(f'{"kernel":58s} ' + ' '.join(f'{c:>8s}' for c in cols))
    total = collections.Counter()
    for (k, c), name in zip(kernels.items(), demangle):
        short = re.sub(r'\(.*', '', name).replace('b200det::', '').replace('void ', '')[:58]
        print(f'{short:58s} ' + ' '.join(f'{c[x]:8d}' for x in cols))
        total.update(c)
    print(f'{"TOTAL":58s} ' + ' '.join(f'{total[x]:8d}' for x in cols))


if __name__ == '__main__':
    main()
