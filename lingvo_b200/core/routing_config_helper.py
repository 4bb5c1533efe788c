"""Routing-Transformer encoder configuration (ref `lingvo/core/routing_config_helper.py`)."""
from lingvo_b200.core import batch_major_attention as bma
from lingvo_b200.core import hyperparams


class RoutingTransformerEncoderParams(hyperparams.Params):
  """Typed bag of the knobs `SetupRoutingTransformerEncoder` consumes (ref :22)."""

  def __init__(self):
    super().__init__()
    self.Define('num_layers', 6, 'Layers.')
    self.Define('model_dim', 512, 'Model dim.')
    self.Define('hidden_dim', 2048, 'FFN hidden dim.')
    self.Define('num_heads', 8, 'Heads.')
    self.Define('attention_window', 256, 'Local attention window of the non-routing heads.')
    self.Define('num_clusters', 16, 'k-means clusters for routing heads.')
    self.Define('dropout_prob', 0.1, 'Dropout.')


def SetupRoutingTransformerEncoder(model_dim, vocab_size, num_layers, num_heads, hidden_dim,
                                   attention_window=256, num_clusters=16, residual_dropout_prob=0.1,
                                   input_dropout_prob=0.0, atten_dropout_prob=0.0,
                                   relu_dropout_prob=0.0, add_unnormalized_residuals=False):
  """Stack whose self-attention is local (band) attention; the routing heads of the
  reference are approximated by the same band restricted to the cluster window."""
  del vocab_size, input_dropout_prob, relu_dropout_prob, num_clusters
  p = bma.StackedTransformerLayers.Params().Set(
      num_layers=num_layers, mdl_dim=model_dim, hidden_dim=hidden_dim,
      num_atten_heads=num_heads, dropout_prob=residual_dropout_prob,
      add_unnormalized_input=add_unnormalized_residuals, final_layer_norm=True)
  tpl = p.transformer_layer_params_tpl
  tpl.tr_atten_tpl.atten_tpl = bma.LocalSelfAttention.Params().Set(
      left_context=attention_window // 2 + 1, right_context=attention_window // 2,
      use_bias=False, enable_per_dim_scale=False, atten_dropout_prob=atten_dropout_prob)
  return p
