"""LXMERT configuration (reference: hg_transformers/configuration_lxmert.py:103-170,
lxmert_config/config.json:1-33).  A plain attribute bag: the hot path only reads sizes."""
import copy
import json


class LxmertConfig:
    model_type = "lxmert"

    def __init__(self, vocab_size=30522, hidden_size=768, ans_num=2486, num_attention_heads=12,
                 intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1,
                 attention_probs_dropout_prob=0.1, max_position_embeddings=512, type_vocab_size=2,
                 initializer_range=0.02, layer_norm_eps=1e-12, l_layers=9, x_layers=5, r_layers=5,
                 visual_feat_dim=2048, visual_pos_dim=4, **kwargs):
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.ans_num = ans_num
        self.num_attention_heads = num_attention_heads
        self.intermediate_size = intermediate_size
        self.hidden_act = hidden_act
        self.hidden_dropout_prob = hidden_dropout_prob
        self.attention_probs_dropout_prob = attention_probs_dropout_prob
        self.max_position_embeddings = max_position_embeddings
        self.type_vocab_size = type_vocab_size
        self.initializer_range = initializer_range
        self.layer_norm_eps = layer_norm_eps
        self.l_layers = l_layers
        self.x_layers = x_layers
        self.r_layers = r_layers
        self.visual_feat_dim = visual_feat_dim
        self.visual_pos_dim = visual_pos_dim
        self.num_hidden_layers = {"vision": r_layers, "cross_encoder": x_layers, "language": l_layers}
        self.output_attentions = False
        self.output_hidden_states = False
        self.pruned_heads = {}
        for k, v in kwargs.items():
            setattr(self, k, v)

    @classmethod
    def from_json_file(cls, path):
        with open(path) as f:
            return cls(**json.load(f))

    def to_dict(self):
        return copy.deepcopy(self.__dict__)
