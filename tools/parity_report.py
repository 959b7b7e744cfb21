"""Collect the measured flip / agreement rates written by tests/test_strict_parity.py (gpurun_out/strict_parity_*.json) into
profiles/r2_strict_parity.md.   python tools/parity_report.py"""
import glob
import json
import os

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = ['# Strict end-to-end parity (round 2): CUDA path, unforced, vs the bf16-emulating oracle',
       '',
       'Source: `tests/test_strict_parity.py` on a B200 (`gpurun_out/strict_parity_*.json`, copied here by',
       '`tools/parity_report.py`).  Same weights, inputs and injected randomness; nothing forced to the oracle\'s decisions.',
       '"c1" = conditioned variant (positive biases in front of the <= 30-pixel InstanceNorms), "c0" = reference init.', '']
for f in sorted(glob.glob(os.path.join(REPO, 'gpurun_out', 'strict_parity_train_*.json'))):
    rows = json.load(open(f))
    out += ['## ' + os.path.basename(f), '',
            '| step | posterior sign flips | sampled bit flips | arg-max agreement | tie-consistent | logits rel | reward abs | '
            'value abs | loss w/o value abs | min grad cos downstream / bottleneck / encoder |', '|---|---|---|---|---|---|---|---|---|---|']
    for r in rows:
        out.append('| %d | %.4g | %.4g | %.5f | %.5f | %.2e | %.2e | %.2e | %.2e | %.4f / %.4f / %.4f |' % (
            r['step'], r['posterior_sign_flip'], r['sampled_bit_flip'], r['argmax_agreement'], r['argmax_tie_consistent'],
            r['logits_rel'], r.get('reward_abs', float('nan')), r['value_abs'], r.get('loss_without_value_abs', float('nan')),
            r.get('grad_min_cos_downstream', float('nan')), r.get('grad_min_cos_at_bottleneck', float('nan')),
            r.get('grad_min_cos_encoder', float('nan'))))
    out.append('')
for f in sorted(glob.glob(os.path.join(REPO, 'gpurun_out', 'strict_parity_rollout_*.json'))):
    rows = json.load(open(f))
    out += ['## ' + os.path.basename(f) + ('  (trained weights, free-running)' if rows[0].get('trained') else
                                          '  (random init, frames re-synchronised every step)'), '',
            '| step | sampled bit flips | new-frame agreement | tie-consistent | reward abs err |', '|---|---|---|---|---|']
    pick = rows if len(rows) <= 6 else rows[:3] + rows[9::10]
    for r in pick:
        out.append('| %d | %.4g | %.5f | %.5f | %.3g |' % (r['step'], r['sampled_bit_flip'], r['new_frame_agreement'],
                                                          r['new_frame_tie_consistent'], r['reward_abs_err']))
    out.append('')
dp = os.path.join(REPO, 'gpurun_out', 'dp_2rank_vs_1rank.txt')
if os.path.exists(dp):
    out += ['## Data parallel: 2 NCCL ranks vs 1 rank on the global batch (tests/test_dp_nccl.py, `gpurun --gpus 2`)', '', '```',
            open(dp).read().strip(), '```', '']
open(os.path.join(REPO, 'profiles', 'r2_strict_parity.md'), 'w').write('\n'.join(out))
print('\n'.join(out)[:3000])
