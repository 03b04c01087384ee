"""Batch formats of the mPLUG VQA loaders (reference mPLUG/dataset/__init__.py:116-135) and a synthetic dataset of the
same item format.  The reference's image datasets, transforms and samplers need the VQA / COCO / VG files (not shipped)."""
import torch
from torch.utils.data import Dataset


def vqa_collate_fn(batch):
    """[(image, question, answers, weights)] -> (images [B,3,H,W], questions, flat answers, flat weights, answers per
    question)."""
    images, questions, answers, weights, n = [], [], [], [], []
    for image, question, answer, weight in batch:
        images.append(image)
        questions.append(question)
        answers += answer
        weights += weight
        n.append(len(answer))
    return torch.stack(images, dim=0), questions, answers, torch.Tensor(weights), n


def vqa_bias_collate_fn(batch):
    """The same with one language-prior bias value per answer appended (VQA-CP debiasing)."""
    images, questions, answers, weights, n, biases = [], [], [], [], [], []
    for image, question, answer, weight, bias in batch:
        images.append(image)
        questions.append(question)
        answers += answer
        weights += weight
        n.append(len(answer))
        biases += bias
    return torch.stack(images, dim=0), questions, answers, torch.Tensor(weights), n, torch.Tensor(biases)


class SyntheticVQAImageDataset(Dataset):
    """Items shaped like the reference's vqa_dataset in training mode: (image [3,res,res], question string, list of
    answer strings, list of answer weights[, list of biases]); text is made of the given vocabulary words."""

    def __init__(self, n, image_res=384, words=("what", "color", "is", "the", "cat", "two", "red", "yes", "no", "dog"),
                 with_bias=True, seed=49, eos="[SEP]"):
        g = torch.Generator().manual_seed(seed)
        self.images = torch.randn(n, 3, image_res, image_res, generator=g)
        pick = lambda k: " ".join(words[int(i)] for i in torch.randint(0, len(words), (k,), generator=g))  # noqa: E731
        self.questions = [pick(int(torch.randint(3, 9, (1,), generator=g))) for _ in range(n)]
        counts = torch.randint(1, 4, (n,), generator=g)
        self.answers = [[pick(int(torch.randint(1, 3, (1,), generator=g))) + eos for _ in range(int(c))] for c in counts]
        self.weights = [[float(w) for w in torch.rand(int(c), generator=g)] for c in counts]
        self.biases = [[float(b) * 0.5 for b in torch.rand(int(c), generator=g)] for c in counts] if with_bias else None

    def __len__(self):
        return len(self.questions)

    def __getitem__(self, i):
        item = (self.images[i], self.questions[i], self.answers[i], self.weights[i])
        return item + (self.biases[i],) if self.biases is not None else item


WORDS = ("what", "color", "is", "the", "cat", "two", "red", "yes", "no", "dog")


class WhitespaceTokenizer:
    """Stand-in for the BERT tokenizer on synthetic text (no vocabulary file ships): whitespace words of a fixed list,
    BERT's special ids ([PAD] 0, [CLS] 101, [SEP] 102 -- the ids the beam search hard-codes), ``padding='longest'``
    batches with ``input_ids`` / ``attention_mask`` and ``.to(device)``, and ``decode``."""
    pad_token_id, cls_token_id, sep_token_id = 0, 101, 102

    def __init__(self, words=WORDS, first_id=103):
        self.ids = {w: first_id + i for i, w in enumerate(words)}
        self.words = {i: w for w, i in self.ids.items()}
        self.words.update({0: "[PAD]", 101: "[CLS]", 102: "[SEP]"})
        self.vocab_size = first_id + len(words)

    class Encoding:
        def __init__(self, input_ids, attention_mask):
            self.input_ids, self.attention_mask = input_ids, attention_mask

        def to(self, device):
            return WhitespaceTokenizer.Encoding(self.input_ids.to(device), self.attention_mask.to(device))

    def __call__(self, texts, padding="longest", truncation=False, max_length=None, return_tensors="pt"):
        rows = []
        for text in texts:
            body = [self.ids[w] for w in text.replace("[SEP]", " ").split()]
            row = [self.cls_token_id] + body + [self.sep_token_id]
            rows.append(row[:max_length] if truncation and max_length else row)
        width = max(len(r) for r in rows)
        ids = torch.tensor([r + [self.pad_token_id] * (width - len(r)) for r in rows])
        return self.Encoding(ids, (ids != self.pad_token_id).long())

    def decode(self, ids):
        return " ".join(self.words.get(int(t), "[UNK]") for t in ids)
