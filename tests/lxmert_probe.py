"""Ad-hoc GPU probe: full LXMERT stage-2 forward/backward through the CUDA path vs the golden file."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'compress-robust-vqa_b200')); sys.path.insert(0, ROOT)
import torch
from prune_debias_VQA import build_stage2, SyntheticVQADataset
from oracle import lxmert_oracle as lxo
dev = torch.device('cuda')
g = torch.load(os.path.join(ROOT, 'tests/golden/full_lxmert.pt'), weights_only=False)
t = time.time()
model, masker, margs = build_stage2(2274, device=dev)
torch.cuda.synchronize(); print('build+patch', time.time() - t)
mods = [(n, m) for n, m in model.named_modules() if hasattr(m, 'threshold')]
bad = [n for n, m in mods if int((m.weight_mask > 1e-2).sum()) != g['kept_init'][n]]
print('modules', len(mods), 'kept mismatch', bad)
batch = lxo.synthetic_batch(32, 2274)
b = {k: v.to(dev) for k, v in batch.items()}
model.eval()
loss, logits, pooled = model(b['ids'], b['feats'], b['pos'], labels=b['target'])
def rel(a, r): return float((a.cpu() - r).abs().max() / r.abs().max())
print('loss', float(loss), float(g['loss_normal']), 'logits rel', rel(logits, g['logits']), 'pooled rel', rel(pooled, g['pooled']))
from crvqa import ops
l2, sc = ops.vqa_loss_bce(logits, b['target'])
print('fused bce', float(l2), 'score', float(sc), float(g['score']))
l3, _ = ops.vqa_loss_lpf(logits, b['bias'], b['max_label'], 5.0, b['target'])
print('fused lpf', float(l3), float(g['loss_lpf']))
l2.backward()
for n in ['lxmert.encoder.layer.0.attention.self.query', 'lxmert.pooler.dense', 'lxmert.embeddings.word_embeddings', 'lxmert.encoder.visn_fc.box_fc', 'lxmert.encoder.x_layers.2.visual_attention.att.query', 'lxmert.encoder.r_layers.3.output.dense']:
    m = dict(mods)[n]
    st = g['grad_stats_normal'][n]
    gr = m.weight_mask.grad
    samp = gr.reshape(-1)[:: max(1, gr.numel() // 512)][:512].cpu()
    print(n, 'l2', float(gr.double().norm()), st['l2'], 'nnz', int((gr != 0).sum()), st['nnz'], 'sample rel', float((samp - st['sample']).abs().max() / st['sample'].abs().max().clamp_min(1e-30)))
print('nograd', [n for n, m in mods if m.weight_mask.grad is None])
