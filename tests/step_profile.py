"""Kernel-time breakdown of one eager training step via torch.profiler (analysis only, not a bench number)."""
import os, sys, collections, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'compress-robust-vqa_b200')); sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from hg_transformers.data.data_collator import TrimCollator
from hg_transformers.mask_trainer_Robust_VQA import Trainer
from hg_transformers.training_args import TrainingArguments
from oracle import lxmert_oracle as lxo
from prune_debias_VQA import build_stage2, init_optimizer
dev = torch.device('cuda')
B, A = 256, 3129
targs = TrainingArguments(output_dir='/tmp/o', per_gpu_train_batch_size=B, logging_steps=100, seed=49, Masker_type='lpf', training_type='Masker', save_steps=0, dataloader_num_workers=0)
model, masker, margs = build_stage2(A, device=dev, seed=49)
opt, sch = init_optimizer(model, targs, B * 10000)
tr = Trainer(model=model, args=targs, model_args=margs, data_collator=TrimCollator(), optimizers=(opt, sch), masker=masker)
tr._setup_engine(opt)
host = lxo.synthetic_batch(B, A)
inputs = [host[k].to(dev) if k else torch.arange(B) for k in ["ids", "feats", "pos", "target", None, None, "bias", "max_label"]]
tr._zero_grad(opt)
for _ in range(3):
    tr._device_step(model, inputs, opt); sch.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr._device_step(model, inputs, opt); sch.step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r'<.*', '', ev.name)[:70]
        agg[name][0] += 1; agg[name][1] += ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time; tot += agg[name][1] * 0
tot = sum(v[1] for v in agg.values())
print('total kernel us', tot)
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
    print(f'{t:9.1f} us {100*t/tot:5.1f}% n={c:4d} avg={t/c:7.1f} {k}')

# per-shape masked-GEMM timings (CUDA events around each launch, eager)
from crvqa import ops
ops.PROFILE = []
tr._device_step(model, inputs, opt); sch.step()
torch.cuda.synchronize()
shape = collections.defaultdict(lambda: [0, 0.0])
for kind, M, N, K, s, e in ops.PROFILE:
    d = shape[(kind, M, N, K)]; d[0] += 1; d[1] += s.elapsed_time(e)
ops.PROFILE = None
print('kind M N K : launches total_ms avg_us TFLOP/s')
for (kind, M, N, K), (c, t) in sorted(shape.items(), key=lambda x: -x[1][1]):
    print(f'{kind:3s} {M:5d} {N:5d} {K:5d} : {c:3d} {t:7.3f} {1e3*t/c:7.1f} {2.0*M*N*K*c/(t*1e-3)/1e12:7.0f}')
