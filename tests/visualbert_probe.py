"""VisualBERT stage-2 mask training (BASELINE config 3: uniform zero rate 0.7, lr 5e-5, batch 256) step time through
the visualBERT trainer: whole-step CUDA graph, fused fast path.  Analysis only (bench.py measures config 2)."""
import logging, os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'compress-robust-vqa_b200')); sys.path.insert(0, ROOT)
import torch
from hg_transformers.data.data_collator import TrimCollator
from hg_transformers.mask_trainer_visualBERT_VQA import Trainer
from hg_transformers.modeling_visualbert import VisualBertForMultipleChoice, visualBERTConfig
from hg_transformers.training_args import TrainingArguments
from masking import maskers_visualBert as mk
from masking import sparsity_control as spc
from prune_debias_VQA import init_optimizer
B, A, T, R = int(os.environ.get('B', 256)), 3129, 20, 36
dev = torch.device('cuda')
torch.manual_seed(49)
cfg = visualBERTConfig(ans_num=A)
targs = TrainingArguments(output_dir='/tmp/vb', per_gpu_train_batch_size=B, logging_steps=100, seed=49, Masker_type='normal',
                          training_type='Masker', save_steps=0, learning_rate=5e-5, dataloader_num_workers=0)
model = VisualBertForMultipleChoice(cfg).to(dev)
log = logging.getLogger('vb'); log.setLevel(logging.ERROR)
conf = types.SimpleNamespace(masking_scheduler_conf_={'final_sparsity': 0.7, 'sparsity_warmup_interval_epoch': 0.1, 'lambdas_lr': 0.0,
                                                      'init_epoch': 0, 'final_epoch': 1}, logger=log, num_epochs=1)
masker = mk.Masker(masker_scheduler=spc.MaskerScheduler(conf), logger=log, mask_biases=False,
                   structured_masking_info={'structured_masking': None, 'structured_masking_types': None, 'force_masking': 'bert'},
                   threshold=1e-2, init_scale=2e-2, which_ptl='visual_bert', controlled_init='magnitude')
masker.patch_modules(model, mk.chain_module_names('visual_bert', list(range(12)), ['K', 'Q', 'V', 'AO', 'I', 'O', 'P', 'E']), 'MaskedLinear1')
n_scores = sum(m.weight_mask.numel() for _, m in model.named_modules() if hasattr(m, 'threshold'))
opt, sch = init_optimizer(model, targs, B * 10000)
tr = Trainer(model=model, args=targs, model_args=types.SimpleNamespace(structured=False), data_collator=TrimCollator(),
             optimizers=(opt, sch), masker=masker)
tr._setup_engine(opt); tr.global_step = 0
g = torch.Generator().manual_seed(49)
ids = torch.randint(1, cfg.vocab_size, (B, T), generator=g)
feats = torch.randn(B, R, cfg.visual_embedding_dim, generator=g)
target = (torch.rand(B, A, generator=g) > 0.999).float() * torch.rand(B, A, generator=g)
bias = torch.rand(B, A, generator=g) * 0.01
inputs = [t.to(dev) for t in (ids, feats, torch.rand(B, R, 4), target, torch.arange(B), torch.arange(B), bias, target.argmax(1))]
graphed = tr._make_graphed_step(model, opt, sch)
tr._zero_grad(opt)
def step():
    if graphed is not None:
        return graphed.step(inputs)[0]
    loss, _ = tr._device_step(model, inputs, opt); sch.step(); return loss
for _ in range(6): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n): loss = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f'visualbert B={B} scores={n_scores} fast_path={model.visual_bert.encoder._fast_plans() is not None} graph={graphed is not None}: {ms:.2f} ms/step, {B/ms*1e3:.0f} samples/s, loss {float(loss):.4f}')
