"""VisualBERT configuration (reference hg_transformers/configuration_visualbert.py:107-150)."""


class VisualBertConfig:
    model_type = "visual_bert"

    def __init__(self, vocab_size=30522, hidden_size=768, ans_num=2274, visual_embedding_dim=512,
                 num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, hidden_act="gelu",
                 hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, max_position_embeddings=512,
                 type_vocab_size=2, initializer_range=0.02, layer_norm_eps=1e-12, bypass_transformer=False,
                 special_visual_initialize=True, pad_token_id=1, bos_token_id=0, eos_token_id=2, **kwargs):
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.ans_num = ans_num
        self.visual_embedding_dim = visual_embedding_dim
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.intermediate_size = intermediate_size
        self.hidden_act = hidden_act
        self.hidden_dropout_prob = hidden_dropout_prob
        self.attention_probs_dropout_prob = attention_probs_dropout_prob
        self.max_position_embeddings = max_position_embeddings
        self.type_vocab_size = type_vocab_size
        self.initializer_range = initializer_range
        self.layer_norm_eps = layer_norm_eps
        self.bypass_transformer = bypass_transformer
        self.special_visual_initialize = special_visual_initialize
        self.pad_token_id, self.bos_token_id, self.eos_token_id = pad_token_id, bos_token_id, eos_token_id
        self.pruned_heads = {}
        for k, v in kwargs.items():
            setattr(self, k, v)


visualBERTConfig = VisualBertConfig  # the drivers use this spelling (prune_debias_VQA_visualBERT.py)
