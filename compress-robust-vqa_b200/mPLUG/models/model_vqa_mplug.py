"""``MPLUG`` -- the VQA network of the reference (mPLUG/models/model_vqa_mplug.py:13-173), training path.

Same constructor arguments, sub-module names (``visual_encoder.visual``, ``text_encoder``, ``fusion_encoder``,
``text_decoder`` and their ``_m`` momentum twins) and ``forward(image, question, answer, alpha, k, weights, train,
bias)`` semantics: CLIP ViT image states -> BERT text encoder -> skip-connected fusion encoder -> [image; question]
states repeated ``k[b]`` times -> causal answer decoder; loss = sum_b weights * nll (optionally (1 - bias) weighted)
/ batch.  With ``config['distill']`` the momentum twins are updated every step as in the reference -- but the
reference never forwards ``alpha`` to the decoder (:96-104), so its distillation term is multiplied by the decoder's
default ``alpha=0`` and the loss is the plain one; the twins' forward pass, whose only product is that zero-weighted
term, is therefore skipped here (``MPLUG.run_unused_distill_forward = True`` runs it anyway).

The reference initialises from checkpoints (``from_pretrained`` of bert-base-uncased, ``ckpts/ViT-B-16.tar``); none
ship, so the stacks are randomly initialised and ``load_state_dict`` of a reference checkpoint works key for key
(except the CLIP text tower, which mPLUG-VQA never runs and this package does not build).  ``train=False`` runs the
beam search of ``predictor.TextGenerator``; the closed-set alternative ``rank_answer`` (:188-245) is built too.
"""
import os

import torch
import torch.nn.functional as F
from torch import nn

from .modeling_mplug import BertConfig, BertLMHeadModel, BertModel, FusionModel
from .predictor import TextGenerator
from .visual_transformers import initialize_clip


class MPLUG(nn.Module):
    run_unused_distill_forward = False
    # Opt-in (CRVQA_MPLUG_BF16_ACTIVATIONS=1): run the whole forward under bf16 autocast on a GPU, which is the
    # reference's DeepSpeed-bf16 arithmetic -- bf16 activations between the masked GEMMs (which then read and write bf16
    # directly), LayerNorm / softmax / loss still in fp32.  Default off: fp32 activations, bf16 only inside the masked
    # GEMMs and the attention cores.
    bf16_activations = os.environ.get("CRVQA_MPLUG_BF16_ACTIVATIONS", "0") == "1"

    def __init__(self, tokenizer=None, config=None):
        super().__init__()
        self.tokenizer = tokenizer
        self.pad_token_id = getattr(tokenizer, "pad_token_id", 0) if tokenizer is not None else 0
        self.module_setting(config)
        self.visual_encoder, _ = initialize_clip(config)
        self.text_encoder = BertModel(self.config_encoder, add_pooling_layer=False)
        self.fusion_encoder = FusionModel(self.config_fusion, add_pooling_layer=False)
        self.text_decoder = BertLMHeadModel(self.config_decoder)
        self.init_distill(config)
        # generation settings: the reference's driver copies --beam_size / --min_length / --max_length into the config
        # (vqa_mplug.py:494-496, defaults 5 / 1 / 10)
        self.beam_generator = TextGenerator(
            {"beam_size": config.get("beam_size", 5), "min_length": config.get("min_length", 1),
             "max_length": config.get("max_length", 10)}, self.text_decoder)

    # -- configuration ---------------------------------------------------------------------------
    @staticmethod
    def _bert_config(config):
        src = config["bert_config"]
        return BertConfig(**src) if isinstance(src, dict) else BertConfig.from_json_file(src)

    def module_setting(self, config):
        self.config_encoder = self._bert_config(config)
        self.config_encoder.num_hidden_layers = self.config_encoder.text_encoder_layers
        self.config_fusion = self._bert_config(config)
        self.config_decoder = self._bert_config(config)
        self.config_decoder.add_cross_attention = True
        self.config_decoder.num_hidden_layers = self.config_decoder.text_decode_layers
        self.large = False
        if self.config_encoder.hidden_size != config["vision_width"]:
            self.visn_fc = nn.Linear(config["vision_width"], self.config_encoder.hidden_size)
            self.visn_layer_norm = nn.LayerNorm(self.config_encoder.hidden_size, eps=1e-12)
            self.dropout = nn.Dropout(self.config_encoder.hidden_dropout_prob)
            self.large = True
        self.use_checkpoint = config.get("use_checkpoint", True)

    def init_distill(self, config):
        self.distill = config["distill"]
        if not self.distill:
            return
        self.visual_encoder_m, _ = initialize_clip(config)
        self.text_encoder_m = BertModel(self.config_encoder, add_pooling_layer=False)
        self.fusion_encoder_m = FusionModel(self.config_fusion, add_pooling_layer=False)
        self.text_decoder_m = BertLMHeadModel(self.config_decoder)
        # the reference pairs exactly these three (the fusion twin is initialised separately and never updated)
        self.model_pairs = [[self.visual_encoder, self.visual_encoder_m], [self.text_encoder, self.text_encoder_m],
                            [self.text_decoder, self.text_decoder_m]]
        if self.large:
            self.visn_fc_m = nn.Linear(config["vision_width"], self.config_encoder.hidden_size)
            self.visn_layer_norm_m = nn.LayerNorm(self.config_encoder.hidden_size, eps=1e-12)
            self.dropout_m = nn.Dropout(self.config_encoder.hidden_dropout_prob)
            self.model_pairs.extend([[self.visn_fc, self.visn_fc_m], [self.visn_layer_norm, self.visn_layer_norm_m]])
        self.copy_params()
        self.momentum = 0.995

    @torch.no_grad()
    def copy_params(self):
        for online, twin in self.model_pairs:
            for p, p_m in zip(online.parameters(), twin.parameters()):
                p_m.data.copy_(p.data)
                p_m.requires_grad = False

    @torch.no_grad()
    def _momentum_update(self):
        """p_m <- p_m * momentum + p * (1 - momentum) for every paired parameter (:152-156), as three multi-tensor
        launches per dtype / device group instead of three launches per parameter; same products, same sum, same
        rounding as the per-parameter expression."""
        online, twins = [], []
        for a, b in self.model_pairs:
            for p, p_m in zip(a.parameters(), b.parameters()):
                online.append(p.data)
                twins.append(p_m.data)
        if not twins:
            return
        if twins[0].is_cuda and os.environ.get("CRVQA_MPLUG_FUSED", "1") != "0":
            # one launch over every pair (crv_momentum_update: 12 B per element, the same two products and one sum)
            from crvqa import ops
            plan = getattr(self, "_momentum_plan", None)
            try:
                if plan is None or plan.key != ops.MomentumPlan.key_of(online, twins):
                    plan = self._momentum_plan = ops.MomentumPlan(online, twins)
            except ValueError:
                plan = None
            if plan is not None:
                plan.run(self.momentum)
                self._drop_twin_caches()
                return
        fresh = torch._foreach_mul(online, 1.0 - self.momentum)
        torch._foreach_mul_(twins, self.momentum)
        torch._foreach_add_(twins, fresh)
        self._drop_twin_caches()

    def _drop_twin_caches(self):
        # the twins' frozen weights and scores just moved underneath their masked modules (through .data, which does not
        # bump the version counters those modules key their bf16 operand caches on): drop the caches
        for _, twin in self.model_pairs:
            for m in twin.modules():
                if hasattr(m, "drop_masked_weight"):
                    m._w16 = None
                    m.drop_masked_weight()

    # -- towers ----------------------------------------------------------------------------------
    def _image_states(self, image, twin=False):
        enc = self.visual_encoder_m if twin else self.visual_encoder
        x = enc.visual(image, skip_last_layer=True, use_checkpoint=False)
        if self.large:
            fc, ln, drop = ((self.visn_fc_m, self.visn_layer_norm_m, self.dropout_m) if twin
                            else (self.visn_fc, self.visn_layer_norm, self.dropout))
            x = drop(ln(fc(x)))
        return x

    def _question_states(self, image_embeds, image_atts, question, twin=False):
        text_enc = self.text_encoder_m if twin else self.text_encoder
        fusion = self.fusion_encoder_m if twin else self.fusion_encoder
        text = text_enc(question.input_ids, attention_mask=question.attention_mask).last_hidden_state
        image_out, question_out = fusion(encoder_embeds=text, attention_mask=question.attention_mask,
                                         encoder_hidden_states=image_embeds, encoder_attention_mask=image_atts)
        return torch.cat([image_out, question_out], 1)

    @staticmethod
    def _repeat(x, k):
        """Row b repeated k[b] times (the reference builds Python lists and stacks them, :54-60)."""
        total = int(sum(int(v) for v in k))          # known on the host: no device sync to size the output
        return x.repeat_interleave(torch.as_tensor(k, device=x.device), dim=0, output_size=total)

    def forward(self, image, question, answer=None, alpha=0, k=None, weights=None, train=True, bias=None):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bool(self.bf16_activations and image.is_cuda)):
            return self._forward(image, question, answer, alpha, k, weights, train, bias)

    def _forward(self, image, question, answer, alpha, k, weights, train, bias):
        image = image.to(dtype=next(self.parameters()).dtype)
        image_embeds = self._image_states(image)
        image_atts = torch.ones(image_embeds.size()[:-1], dtype=torch.long, device=image.device)
        if not train:
            states = self._question_states(image_embeds, image_atts, question)
            return self.generation(states, torch.cat([image_atts, question.attention_mask], 1))

        answer_targets = answer.input_ids.masked_fill(answer.input_ids == self.pad_token_id, -100)
        question_states = self._repeat(self._question_states(image_embeds, image_atts, question), k)
        question_atts = self._repeat(torch.cat([image_atts, question.attention_mask], 1), k)

        soft_labels = None
        if self.distill:
            self._momentum_update()
        if self.distill and self.run_unused_distill_forward:
            with torch.no_grad():
                image_embeds_m = self._image_states(image, twin=True)
                states_m = self._repeat(self._question_states(image_embeds_m, image_atts, question, twin=True), k)
                logits_m = self.text_decoder_m(answer.input_ids, attention_mask=answer.attention_mask,
                                               encoder_hidden_states=states_m, encoder_attention_mask=question_atts,
                                               return_logits=True)
                soft_labels = F.softmax(logits_m, dim=-1)
        out = self.text_decoder(answer.input_ids, attention_mask=answer.attention_mask,
                                encoder_hidden_states=question_states, encoder_attention_mask=question_atts,
                                labels=answer_targets, return_dict=True, soft_labels=soft_labels, reduction="none")
        loss = weights * out.loss
        if bias is not None:
            loss = (1 - bias) * loss
        return loss.sum() / image.size(0)

    def generation(self, question_states, question_atts):
        """Beam search over the answer decoder (:181-184): per question a list of token-id tensors and their scores."""
        return self.beam_generator.translate_batch([question_states, question_atts])

    # -- closed-set inference --------------------------------------------------------------------
    @torch.no_grad()
    def encode_question(self, image, question):
        """[image; question] states and their attention mask, as the ``train=False`` branch builds them (:122-134)."""
        image = image.to(dtype=next(self.parameters()).dtype)
        image_embeds = self._image_states(image)
        image_atts = torch.ones(image_embeds.size()[:-1], dtype=torch.long, device=image.device)
        states = self._question_states(image_embeds, image_atts, question)
        return states, torch.cat([image_atts, question.attention_mask], 1)

    def rank_answer(self, question_states, question_atts, answer_ids, answer_atts, k):
        """Rank a fixed candidate list (:188-245): the k candidates whose FIRST token is most probable after [BOS] are
        re-scored by the full sequence log-likelihood; returns (topk_ids [Q, k] into the candidate list, topk_probs)."""
        num_ques = question_states.size(0)
        start_ids = answer_ids[0, 0].repeat(num_ques, 1)                       # the shared [BOS] token
        start = self.text_decoder(start_ids, encoder_hidden_states=question_states,
                                  encoder_attention_mask=question_atts, return_dict=True, reduction="none")
        first_token_probs = F.softmax(start.logits[:, 0, :], dim=1).index_select(dim=1, index=answer_ids[:, 1])
        topk_probs, topk_ids = first_token_probs.topk(k, dim=1)

        input_ids = answer_ids.index_select(0, topk_ids.reshape(-1))           # [Q * k, L], question-major
        input_atts = answer_atts.index_select(0, topk_ids.reshape(-1))
        targets = input_ids.masked_fill(input_ids == self.pad_token_id, -100)
        out = self.text_decoder(input_ids, attention_mask=input_atts,
                                encoder_hidden_states=tile(question_states, 0, k),
                                encoder_attention_mask=tile(question_atts, 0, k),
                                labels=targets, return_dict=True, reduction="none")
        answer_loss = out.loss.view(input_ids.size(0), -1)
        log_probs = torch.cat([topk_probs.view(-1, 1).log(), -answer_loss], dim=1).sum(1).view(num_ques, k)
        topk_probs, rerank = F.softmax(log_probs, dim=-1).topk(k, dim=1)
        return torch.gather(topk_ids, 1, rerank), topk_probs


def tile(x, dim, n_tile):
    """Each slice along ``dim`` repeated ``n_tile`` times in place: [a, b] -> [a, a, b, b] (:247-253)."""
    return x.repeat_interleave(n_tile, dim=dim)
