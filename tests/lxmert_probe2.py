import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'compress-robust-vqa_b200')); sys.path.insert(0, ROOT)
import torch
from prune_debias_VQA import build_stage2
from oracle import lxmert_oracle as lxo
from crvqa import ops
dev = torch.device('cuda')
model, masker, margs = build_stage2(2274, device=dev)
model.eval()
mods = [(n, m) for n, m in model.named_modules() if hasattr(m, 'threshold')]
batch = lxo.synthetic_batch(32, 2274)
b = {k: v.to(dev) for k, v in batch.items()}
res = {}
for kind in ['normal', 'lpf']:
    model.zero_grad()
    _, logits, pooled = model(b['ids'], b['feats'], b['pos'], labels=b['target'])
    loss = ops.vqa_loss_bce(logits, b['target'])[0] if kind == 'normal' else ops.vqa_loss_lpf(logits, b['bias'], b['max_label'], 5.0, b['target'])[0]
    loss.backward()
    gpu = {n: (m.weight_mask.grad.cpu().clone() if m.weight_mask.grad is not None else None) for n, m in mods}
    for operand in ['bf16', 'fp32']:
        params = {k: v.detach().cpu().clone() for k, v in model.state_dict().items() if 'weight_mask' not in k}
        for k in params: params[k].requires_grad_(k.startswith('classifier.'))
        scores = {n: m.weight_mask.detach().cpu().clone().requires_grad_(True) for n, m in mods}
        thr = {n: float(m.threshold) for n, m in mods}
        c = lxo.Ctx(params, scores, thr, operand=operand)
        ref = lxo.training_step(c, batch, kind)
        print(kind, operand, 'loss', float(loss), float(ref['loss']), 'logits', float((logits.detach().cpu()-ref['logits']).abs().max()/ref['logits'].abs().max()))
        worst = []
        for (n, m), g in zip(mods, ref['grads']):
            if gpu[n] is None: continue
            rel = float((gpu[n]-g).double().norm()/(g.double().norm()+1e-30))
            worst.append((rel, n))
        worst.sort(reverse=True)
        print('   worst grads', [(round(r,5), n.replace('lxmert.','')) for r, n in worst[:6]])
        print('   median', sorted(r for r,_ in worst)[len(worst)//2])
