"""Regenerates the measured tables of DESIGN.md from bench.py's own JSON lines (so the document
quotes what the bench printed, not hand-copied numbers):

    python tools/design_tables.py profiles/r02_bench_line.json [profiles/r02_scale_n{1,2,4,8}.json ...]

Replaces the text between `<!-- BEGIN:name -->` / `<!-- END:name -->` markers in DESIGN.md."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def kernel_table(d):
    """Two blocks: the step as the library runs it (sweep hand-over: ONE read of cls per step) and the
    same loop with the hand-over off (`separate_sweeps`: every call sweeps for itself)."""
    B = d['config']['images_per_gpu']
    n, c = d['config']['rows_per_image'], 80
    sweep = B * 4 * n * c
    peak = d['roofline']['peak']
    bound = {'focal_loss': 'HBM', 'score_argmax': 'HBM', 'assign': 'ALU (issue)', 'sparse_losses': 'latency / issue',
             'loss_reduce': '—', 'select_decode_nms': 'latency (L2-resident keys)'}
    head = ['| Kernel (1 launch / step) | Bound | Alg. bytes / launch | Time (CUDA events in the timed region) | Achieved |',
            '|---|---|---|---|---|']

    def block(kernels, handed, ms_step, step_bytes, label):
        names = {'focal_loss': '`focal_all_kernel<4,γ=2>` (beside it on the helper stream: assignment + sparse losses)',
                 'assign': '`retina_assign_tile_kernel` (helper stream, inside the sweep\'s window)',
                 'sparse_losses': '`sparse_loss_kernel` (helper stream, inside the sweep\'s window)',
                 'loss_reduce': '`loss_reduce` (+ finish / peer exchange)',
                 'score_argmax': '`fused_rows_tma_kernel` (focal sum + decoder keys from ONE read of cls; beside it '
                                 'on the helper stream: assignment + sparse losses)' if handed else '`score_argmax_kernel<4>`',
                 'select_decode_nms': '`select_nms_kernel` (cluster per image)'}
        rows = []
        for k, ms in kernels.items():
            a = sweep if k in ('focal_loss', 'score_argmax') else None
            ach = f'{a / (ms / 1e3) / 1e12:.2f} TB/s = **{a / (ms / 1e3) / 1e9 / peak:.2f} × measured peak**' if a else '—'
            rows.append((f'| {names.get(k, k)} | {bound.get(k, "—")} | {a / 1e9:.3f} GB |' if a else
                         f'| {names.get(k, k)} | {bound.get(k, "—")} | — |') + f' {ms:.4f} ms | {ach} |')
        gbs = step_bytes / (ms_step / 1e3) / 1e9
        rows.append(f'| **step: {label}** | HBM | {step_bytes / 1e9:.2f} GB | **{ms_step:.3f} ms** | '
                    f'**{gbs / 1e3:.2f} TB/s = {gbs / peak:.3f} × measured peak** |')
        return rows

    full = B * d['config']['algorithmic_bytes_per_image']
    handed = 'focal_loss' not in d['kernels_ms']
    rows = list(head)
    rows += block(d['kernels_ms'], handed, d['ms_per_step'], full - sweep if handed else full,
                  'criterion + decoder, 2 calls, cls read once (hand-over)' if handed else 'criterion + decoder, 2 calls')
    sep = d.get('separate_sweeps')
    if handed and sep:
        rows.append('| *the same loop with `B200DET_HANDOFF=0` (every call sweeps for itself):* | | | | |')
        rows += block(sep['kernels_ms'], False, sep['ms_per_step'], full, 'criterion + decoder, 2 calls, cls read twice')
    return '\n'.join(rows)


def configs_table(d):
    rows = ['| Config | loss fwd | loss fwd+bwd | decode+NMS | eval step: both calls, cls read once | reference on the host CPU (bounded sample) |', '|---|---|---|---|---|---|']
    for c in d.get('configs', []):
        def cell(k):
            if k not in c:
                return '—'
            v = c[k]
            return f'{v["ms"]:.3f} ms = {v["frac_of_hbm_peak"]:.2f}' if v['frac_of_hbm_peak'] >= 0.3 else f'{v["ms"]:.3f} ms (latency-bound, {v["frac_of_hbm_peak"]:.2f})'
        cpu = c.get('cpu_baseline')
        rows.append(f'| {c["name"]} | {cell("loss_fwd")} | {cell("loss_fwd_bwd")} | {cell("decode_nms")} | {cell("eval_step")} | '
                    + (f'{cpu["value"]:.1f} images/s ({cpu["sample"].split(", stages")[0]}, {cpu["cores"]} cores)' if cpu else '—') + ' |')
    return '\n'.join(rows)


def scale_table(lines):
    rows = ['| N GPUs | images / GPU | ms / step | images/s | speed-up | efficiency | exchange | sharded == unsharded |', '|---|---|---|---|---|---|---|---|']
    base = None
    for d in sorted(lines, key=lambda x: x['n_gpus']):
        if base is None:
            base = d['value'] / d['n_gpus']
        pc = d.get('parity_check')
        rows.append(f'| {d["n_gpus"]} | {d["config"]["images_per_gpu"]} | {d["ms_per_step"]:.4f} | {d["value"]:,.0f} | '
                    f'{d["value"] / base:.2f}× | {d["value"] / base / d["n_gpus"]:.3f} | {d["exchange"].split(" per step")[0] if d["n_gpus"] > 1 else "—"} | '
                    + (f'rel {pc["sharded_vs_unsharded_rel"]:.1e}, labels {"equal" if pc["labels_equal"] else "DIFFER"}' if pc else '—') + ' |')
    return '\n'.join(rows)


def main():
    args = sys.argv[1:]
    only = None
    if args and args[0].startswith('--only='):
        only = args.pop(0).split('=', 1)[1].split(',')
    lines = [load(p) for p in args]
    main_line = next(d for d in lines if d['n_gpus'] == 1)
    blocks = {'kernels': kernel_table(main_line), 'configs': configs_table(main_line)}
    if len(lines) > 1:
        blocks['scaling'] = scale_table(lines)
    path = os.path.join(ROOT, 'DESIGN.md')
    text = open(path).read()
    for name, body in blocks.items():
        if only and name not in only:
            continue
        pat = re.compile(rf'(<!-- BEGIN:{name} -->\n).*?(<!-- END:{name} -->)', re.S)
        if not pat.search(text):
            print(f'marker {name} not found', file=sys.stderr)
            continue
        text = pat.sub(lambda m: m.group(1) + body + '\n' + m.group(2), text)
    open(path, 'w').write(text)


if __name__ == '__main__':
    main()
