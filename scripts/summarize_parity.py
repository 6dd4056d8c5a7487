"""profiles/<tag>_parity.md from the parity_*.jsonl files a `pytest -m gpu` run leaves in gpurun_out/."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'
ops = [json.loads(l) for l in open(os.path.join(ROOT, 'gpurun_out', 'parity_ops.jsonl'))]
mods = [json.loads(l) for l in open(os.path.join(ROOT, 'gpurun_out', 'parity_models.jsonl'))]
out = ['# Parity summary (B200, from gpurun_out/parity_*.jsonl of the full `pytest -m gpu` run)\n',
       'Max-norm relative errors against the float64 CPU oracle unless noted; mode 0 = SE_MODE_F32, 1 = SE_MODE_TF32 (single pass), 2 = SE_MODE_TF32X3 (error-compensated: the benchmarked mode).\n',
       '## Training steps (tests/test_gpu_models.py; cases ending in -x3 run SE_MODE_TF32X3, the others SE_MODE_F32; identical fp32-rounded weights on both sides)\n',
       'Gradient columns: relative L2 over all parameters / worst tensor.  "f32 oracle" = the same step by the oracle in '
       'float32, "flip quantum" = the float64 oracle with the masks of its fragile ReLU inputs (|x| < 4e-6) inverted: '
       'the gradient of a ReLU network moves by that much when fp32 rounding changes the sign of a near-zero pre-activation.\n',
       '| case | step | loss | embeddings | grad norm | grads global | grads worst | f32 oracle | flip quantum | weights after step |',
       '|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|']
for d in mods:
    if d['test'] == 'train_step':
        out.append('| %s | %d | %.1e | %.1e | %.1e | %.1e | %.1e | %.1e | %.1e | %.1e |' % (
            d['case'], d['step'], d['loss'], d['emb'], d['gnorm'], d['grad_global'], d['grad_worst'],
            d['grad_floor_f32_oracle'], d.get('relu_flip_quantum', 0.0), d['weights']))
out.append('\n## Forward vs the reference-graph fixtures (tests/golden/arch_*.npz)\n')
out.append('| architecture | raw output | l2-normalised embedding | classifier prob |\n|---|---:|---:|---:|')
for d in mods:
    if d['test'] == 'forward_fixture':
        out.append('| %s | %.1e | %.1e | %.1e |' % (d['tag'], d['z'], d['emb'], d['prob']))
for d in mods:
    if d['test'] == 'resnet50_step':
        out.append('\nResNet-50 (64x64, B=2) step: loss %.1e, emb %.1e, grads %.1e\n' % (d['loss'], d['emb'], d['grad_global']))
    if d['test'] == 'fast_mode':
        out.append('\nFast mode (SE_MODE_TF32, ResNet-110-fc, B=8, one step): embeddings %.1e, loss %.1e '
                   '(TF32 operands: 10-bit mantissa; does not meet 1e-4 by design)\n' % (d['emb'], d['loss']))
out.append('## Convolution kernels (tests/test_gpu_ops.py)\n')
out.append('| case (N,H,W,Cin,Cout,k,stride,pad,bias) | mode | y | dx | dw | db |\n|---|---:|---:|---:|---:|---:|')
for d in ops:
    if d['test'] == 'conv':
        out.append('| %s | %d | %.1e | %.1e | %.1e | %.1e |' % (d['case'], d['mode'], d['y'], d['dx'], d['dw'], d['db']))
out.append('\n## Fused conv + BatchNorm (`se_conv_bn_fwd`, SE_MODE_TF32) vs the oracle; identical to the two-kernel path\n')
out.append('| case (N,H,W,Cin,Cout,bias,conv relu,residual,bn relu,fused) | launches | conv y | bn out | mean | invstd | moving mean | moving var |\n|---|---:|---:|---:|---:|---:|---:|---:|')
for d in ops:
    if d['test'] == 'conv_bn_fused':
        out.append('| %s | %d | %.1e | %.1e | %.1e | %.1e | %.1e | %.1e |' % (d['case'], d['launches'], d['y'], d['z'], d['mean'],
                                                                             d['invstd'], d['mm'], d['mv']))
out.append('\n## BatchNorm kernels (fp32)\n')
out.append('| case | y | mean | invstd | dx | dgamma | dbeta | dres |\n|---|---:|---:|---:|---:|---:|---:|---:|')
for d in ops:
    if d['test'] == 'bn':
        out.append('| %s | %.1e | %.1e | %.1e | %.1e | %.1e | %.1e | %.1e |' % (d['case'], d['y'], d['mean'], d['invstd'], d['dx'],
                                                                              d['dgamma'], d['dbeta'], d['dres']))
out.append('\n## Retrieval kernel vs float64 and vs the reference rankings (N=256 fixture)\n')
out.append('| fixture | mode | max abs err | value scale | rank mismatches (all positions) | positions with unambiguous float64 order |\n|---|---:|---:|---:|---:|---:|')
for d in ops:
    if d['test'] == 'pairwise':
        out.append('| %s | %d | %.1e | %.3g | %.1e | %.4f |' % (d['key'], d['mode'], d['max_abs_err'], d['scale'], d['rank_mismatch'],
                                                             d['unambiguous']))
out.append('\nEvery mismatching rank position lies where the float64 distances of neighbouring ranks differ by less than 4x the '
           'kernel error (asserted by the test).\n')
with open(os.path.join(ROOT, 'profiles', '%s_parity.md' % tag), 'w') as f:
    f.write('\n'.join(out))
print('\n'.join(out[:24]))
