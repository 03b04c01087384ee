"""Stage-3 (FT_trainedMask, LMH) step time at the stage-2 bench shape (B=256, A=3129): analysis only."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'compress-robust-vqa_b200')); sys.path.insert(0, ROOT)
import torch
import run_vqa_stage3 as s3
from crvqa import ops
from hg_transformers.data.data_collator import TrimCollator
from hg_transformers.mask_trainer_VQA import Trainer
from hg_transformers.modeling_lxmert import LxmertConfig, LxmertForMultipleChoice
from hg_transformers.training_args import TrainingArguments
from oracle import lxmert_oracle as lxo
from prune_debias_VQA import ModelArguments
B, A = int(os.environ.get('B', 256)), 3129
torch.manual_seed(49)
model = LxmertForMultipleChoice(LxmertConfig(ans_num=A)).cuda()
mods = dict(model.lxmert.named_modules()); names = s3.trained_mask_module_names()
ws = [mods[n].weight.detach() for n in names]
thr = ops.kth_value_batched(ws, [max(1, int(w.numel() * 0.7)) for w in ws], use_abs=True)
s3.pruning_model_with_mask(model.lxmert, {f'lxmert.{n}.weight_mask': (w.abs() > thr[i]) for i, (n, w) in enumerate(zip(names, ws))}, 'lxmert')
targs = TrainingArguments(output_dir='/tmp/s3', per_gpu_train_batch_size=B, logging_steps=1000, seed=49, training_type='FT_trainedMask', FT_type='lmh', save_steps=0, dataloader_num_workers=0)
opt, sch = s3.init_optimizer(model, targs, B * 1000)
tr = Trainer(model=model, args=targs, model_args=ModelArguments(), data_collator=TrimCollator(), optimizers=(opt, sch), masker=None)
host = lxo.synthetic_batch(B, A)
inputs = [host[k].cuda() if k else torch.arange(B) for k in ['ids', 'feats', 'pos', 'target', None, None, 'bias', 'max_label']]
model.train()
def step():
    opt.zero_grad(set_to_none=True)
    loss, _ = tr._training_step(model, inputs, opt)
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step(); sch.step()
    return loss
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
t0 = time.perf_counter(); e0.record()
for _ in range(n): loss = step()
e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f'stage3 B={B}: {e0.elapsed_time(e1)/n:.1f} ms/step (device), {(t1-t0)*1e3/n:.1f} ms/step (wall), {B*n/(t1-t0):.0f} samples/s, loss {float(loss):.4f}')
if os.environ.get('PROFILE'):
    import collections, re
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step(); torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = re.sub(r'<.*', '', ev.name)[:70]
            agg[name][0] += 1; agg[name][1] += ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time
    tot = sum(v[1] for v in agg.values())
    print('total kernel us', tot)
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:25]:
        print(f'{t:9.1f} us {100*t/tot:5.1f}% n={c:4d} avg={t/c:7.1f} {k}')
