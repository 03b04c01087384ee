"""Beam-search answer generation of mPLUG-VQA (reference mPLUG/models/predictor.py:33-310, ``TextGenerator`` with
``translate_batch`` -> ``_fast_translate_batch``; the sampling / SCST variants of that file are not built).

Same search as the reference, step for step, so the returned sequences and scores are the reference's:

* every question starts from ``[CLS]`` (id 101) on ``beam_size`` beams, the first beam holding all the probability;
* a step re-runs the decoder on the whole alive prefix (no key/value cache, as the reference), takes
  ``log(softmax(logits))`` of the last position, forbids ``[SEP]`` (id 102) before ``min_length``, adds the beam's running
  log-probability, divides by the length penalty ((5 + t) / 6) ** 0.6 and keeps the ``beam_size`` best (beam, token)
  pairs per question;
* a beam that emits ``[SEP]`` (or any beam at ``max_length``) is recorded as a hypothesis with its penalised score; a
  question is closed when its BEST beam finishes -- all its beams are then recorded, the hypotheses sorted by score
  and the best ``beam_size`` kept; finished-but-not-best beams are NOT removed from the search (the reference keeps
  expanding them), closed questions are dropped from the batch.
"""
import torch


def tile(x, count, dim=0):
    """Each slice along ``dim`` repeated ``count`` times in place (reference :500-519)."""
    return x.repeat_interleave(count, dim=dim)


class TextGenerator(object):
    def __init__(self, args, model, vocab=None, symbols=None, global_scorer=None, logger=None, dump_beam=""):
        self.alpha = 0.6
        self.logger = logger
        self.args = args
        self.model = model
        self.vocab = vocab
        self.symbols = symbols
        self.start_token = 101          # [CLS]
        self.end_token = 102            # [SEP]
        self.global_scorer = global_scorer
        self.beam_size = args["beam_size"]
        self.min_length = args["min_length"]
        self.max_length = args["max_length"]
        self.dump_beam = dump_beam

    def translate_batch(self, encoder_inputs, do_sample=False, out_size=1):
        if do_sample:
            raise NotImplementedError("sampling (top-k / top-p) generation is not built; beam search only")
        with torch.no_grad():
            return self._fast_translate_batch(encoder_inputs, self.max_length, min_length=self.min_length,
                                              out_size=out_size)

    def _fast_translate_batch(self, encoder_inputs, max_length, min_length=0, do_sample=False, out_size=1):
        assert not self.dump_beam and not do_sample
        if len(encoder_inputs) == 3:
            states, mask, prefix = encoder_inputs
        else:
            (states, mask), prefix = encoder_inputs, None
        device = states.device
        beams = self.beam_size
        batch = states.size(0)
        states, mask = tile(states, beams), tile(mask, beams)
        owner = torch.arange(batch, dtype=torch.long, device=device)             # original question of each alive row
        first_row = torch.arange(0, batch * beams, step=beams, dtype=torch.long, device=device)
        if prefix is not None:
            alive = tile(prefix, beams)
        else:
            alive = torch.full([batch * beams, 1], self.start_token, dtype=torch.long, device=device)
        running = torch.tensor([0.0] + [float("-inf")] * (beams - 1), device=device).repeat(batch)
        hypotheses = [[] for _ in range(batch)]
        best_scores = [[] for _ in range(batch)]
        best_preds = [[] for _ in range(batch)]

        for step in range(max_length):
            logits = self.model(alive, encoder_hidden_states=states, encoder_attention_mask=mask, return_dict=True,
                                reduction="none").logits[:, -1, :]
            vocab = logits.size(-1)
            log_probs = torch.log(torch.softmax(logits.view(-1, vocab), dim=-1))
            if step < min_length:
                log_probs[:, self.end_token] = -1e20
            penalty = ((5.0 + (step + 1)) / 6.0) ** self.alpha
            log_probs += running.view(-1).unsqueeze(1)
            scores, flat = (log_probs / penalty).reshape(-1, beams * vocab).topk(beams, dim=-1)
            running = scores * penalty
            from_beam, token = flat // vocab, flat.fmod(vocab)
            rows = from_beam + first_row[:from_beam.size(0)].unsqueeze(1)         # row of the parent beam, old layout
            alive = torch.cat([alive.index_select(0, rows.view(-1)), token.view(-1, 1)], -1)

            finished = token.eq(self.end_token)
            if step + 1 == max_length:
                finished.fill_(1)
            closed = finished[:, 0].eq(1)                                          # the best beam of the question ended
            if finished.any():
                grouped = alive.view(-1, beams, alive.size(-1))
                for i in range(finished.size(0)):
                    q = int(owner[i])
                    if closed[i]:
                        finished[i].fill_(1)
                    for j in finished[i].nonzero().view(-1):
                        hypotheses[q].append((scores[i, j], grouped[i, j, 0:]))
                    if closed[i]:
                        ranked = sorted(hypotheses[q], key=lambda h: float(h[0]), reverse=True)
                        for score, pred in ranked[:beams]:
                            best_scores[q].append(score)
                            best_preds[q].append(pred)
                still = closed.eq(0).nonzero().view(-1)
                if len(still) == 0:
                    break
                running = running.index_select(0, still)
                rows = rows.index_select(0, still)
                owner = owner.index_select(0, still)
                alive = grouped.index_select(0, still).view(-1, alive.size(-1))
            states = states.index_select(0, rows.view(-1))
            mask = mask.index_select(0, rows.view(-1))
        return [p[:out_size] for p in best_preds], [s[:out_size] for s in best_scores]
