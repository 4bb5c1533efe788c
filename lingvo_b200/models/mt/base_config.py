"""Shared MT configuration helpers (ref `lingvo/tasks/mt/base_config.py`)."""

from __future__ import annotations

from lingvo_b200.core import attention
from lingvo_b200.core import layers
from lingvo_b200.core import optimizer
from lingvo_b200.core import py_utils
from lingvo_b200.core import rnn_cell
from lingvo_b200.core import schedule
from lingvo_b200.models.mt import decoder
from lingvo_b200.models.mt import encoder


def InitTrainDatasetParams(vocab_size=None, params=None):
  """Bucketing for RNMT-style training data (ref :31)."""
  p = params
  p.file_random_seed = 0
  p.file_parallelism = 16
  p.file_buffer_size = 10000000
  p.bucket_upper_bound = [10, 14, 19, 26, 36, 50, 70, 98]
  p.bucket_batch_limit = [128] * 8
  if vocab_size is not None:
    p.tokenizer.vocab_size = vocab_size
  return p


def InitTestDatasetParams(vocab_size=None, params=None):
  """Deterministic eval bucketing (ref :80)."""
  p = params
  p.file_random_seed = 27182818
  p.file_parallelism = 1
  p.file_buffer_size = 1
  p.bucket_upper_bound = [10, 14, 19, 26, 36, 50, 70, 98, 137, 200]
  p.bucket_batch_limit = [16] * 8 + [4] * 2
  if vocab_size is not None:
    p.tokenizer.vocab_size = vocab_size
  return p


def InitTransformerTestBuckets(params):
  params.bucket_upper_bound = [10, 14, 19, 26, 36, 50, 70, 98, 137, 200]
  params.bucket_batch_limit = [16] * 8 + [4] * 2
  return params


def InitTransformerTrainBuckets(params):
  params.bucket_upper_bound = [8, 12, 16, 24, 32, 48, 64, 96]
  params.bucket_batch_limit = [512, 341, 256, 170, 128, 85, 64, 42]
  return params


def SetupTransformerEncoder(model_dim, vocab_size, num_layers, num_heads, hidden_dim,
                            residual_dropout_prob=0.1, input_dropout_prob=0.0,
                            atten_dropout_prob=0.0, relu_dropout_prob=0.0,
                            add_unnormalized_residuals=False, packed_input=False):
  """Transformer encoder params (ref :295)."""
  p = encoder.TransformerEncoder.Params()
  p.name = 'enc'
  p.model_dim = model_dim
  p.packed_input = packed_input
  p.token_emb.Set(vocab_size=vocab_size, embedding_dim=model_dim,
                  params_init=py_utils.WeightInit.Gaussian(1.0 / model_dim ** 0.5))
  p.position_emb.Set(embedding_dim=model_dim)
  p.input_dropout_prob = input_dropout_prob
  p.transformer_stack.Set(num_layers=num_layers, mdl_dim=model_dim, hidden_dim=hidden_dim,
                          num_atten_heads=num_heads, dropout_prob=residual_dropout_prob,
                          add_unnormalized_input=add_unnormalized_residuals)
  del atten_dropout_prob, relu_dropout_prob
  return p


def SetupTransformerDecoder(model_dim, vocab_size, num_layers, num_heads, hidden_dim,
                            residual_dropout_prob=0.1, input_dropout_prob=0.0,
                            atten_dropout_prob=0.0, relu_dropout_prob=0.0,
                            label_smoothing_uncertainty=0.1, add_unnormalized_residuals=False,
                            packed_input=False):
  """Transformer decoder params (ref :214)."""
  p = decoder.TransformerDecoder.Params()
  p.name = 'dec'
  p.source_dim = model_dim
  p.model_dim = model_dim
  p.num_trans_layers = num_layers
  p.num_atten_heads = num_heads
  p.hidden_dim = hidden_dim
  p.packed_input = packed_input
  p.residual_dropout_prob = residual_dropout_prob
  p.input_dropout_prob = input_dropout_prob
  p.token_emb.Set(vocab_size=vocab_size, embedding_dim=model_dim,
                  params_init=py_utils.WeightInit.Gaussian(1.0 / model_dim ** 0.5))
  p.position_emb.Set(embedding_dim=model_dim)
  p.softmax.Set(num_classes=vocab_size, input_dim=model_dim)
  p.per_word_avg_loss = True
  if label_smoothing_uncertainty:
    p.label_smoothing = layers.UniformLabelSmoother.Params().Set(
        num_classes=vocab_size, uncertainty=label_smoothing_uncertainty)
  p.target_seq_len = 300
  p.beam_search.length_normalization = 0.5
  p.beam_search.coverage_penalty = 0.0
  del atten_dropout_prob, relu_dropout_prob, add_unnormalized_residuals
  return p


def SetupTransformerParams(p, name, vocab_size, model_dim, hidden_dim, num_heads, num_layers,
                           learning_rate, warmup_steps, residual_dropout_prob=0.1,
                           input_dropout_prob=0.0, atten_dropout_prob=0.0,
                           relu_dropout_prob=0.0, label_smoothing_uncertainty=0.1,
                           is_transparent=False, activation='RELU',
                           add_unnormalized_residuals=False, atten_hidden_dim=0,
                           packed_input=False):
  """Fills a `TransformerModel.Params()` (ref :127)."""
  del is_transparent, activation, atten_hidden_dim
  p.name = name
  p.encoder = SetupTransformerEncoder(
      model_dim, vocab_size, num_layers, num_heads, hidden_dim, residual_dropout_prob,
      input_dropout_prob, atten_dropout_prob, relu_dropout_prob, add_unnormalized_residuals,
      packed_input)
  p.decoder = SetupTransformerDecoder(
      model_dim, vocab_size, num_layers, num_heads, hidden_dim, residual_dropout_prob,
      input_dropout_prob, atten_dropout_prob, relu_dropout_prob, label_smoothing_uncertainty,
      add_unnormalized_residuals, packed_input)
  p.train.Set(learning_rate=learning_rate, optimizer=optimizer.Adam.ParamsB(),
              clip_gradient_norm_to_value=0.0, grad_norm_to_clip_to_zero=0.0,
              lr_schedule=schedule.TransformerSchedule.Params().Set(
                  warmup_steps=warmup_steps, worker_replicas=1, model_dim=model_dim))
  p.eval.samples_per_summary = 12000
  return p


def SetupRNMTParams(p, name, vocab_size, embedding_dim, hidden_dim, num_heads,
                    num_encoder_layers, num_decoder_layers, learning_rate, l2_regularizer_weight,
                    lr_warmup_steps, lr_decay_start, lr_decay_end, lr_min, ls_uncertainty,
                    atten_dropout_prob, residual_dropout_prob, adam_beta2, adam_epsilon,
                    add_summary=True):
  """Fills an `RNMTModel.Params()` (ref :384)."""
  del add_summary
  p.name = name
  default_init = py_utils.WeightInit.Uniform(0.04)
  emb_init = py_utils.WeightInit.Gaussian(0.01)
  p.encoder = encoder.MTEncoderBiRNN.Params().Set(
      name='enc', num_lstm_layers=num_encoder_layers, lstm_cell_size=hidden_dim,
      encoder_out_dim=hidden_dim, dropout_prob=residual_dropout_prob)
  p.encoder.emb.Set(vocab_size=vocab_size, embedding_dim=embedding_dim, params_init=emb_init)
  p.encoder.lstm_tpl.params_init = default_init
  d = decoder.MTDecoderV1.Params().Set(
      name='dec', source_dim=hidden_dim, rnn_cell_dim=hidden_dim,
      rnn_layers=num_decoder_layers, dropout_prob=residual_dropout_prob,
      feed_attention_context_vec_to_softmax=True, per_word_avg_loss=True)
  d.emb.Set(vocab_size=vocab_size, embedding_dim=embedding_dim, params_init=emb_init)
  d.attention = attention.MultiHeadedAttention.Params().Set(
      hidden_dim=hidden_dim, num_attention_heads=num_heads, context_dim=hidden_dim,
      atten_dropout_prob=atten_dropout_prob, params_init=default_init)
  d.atten_rnn_cell_tpl = rnn_cell.LayerNormalizedLSTMCellSimple.Params().Set(
      params_init=default_init)
  d.rnn_cell_tpl = rnn_cell.LayerNormalizedLSTMCellSimple.Params().Set(params_init=default_init)
  d.softmax.Set(num_classes=vocab_size)
  d.label_smoothing = layers.UniformLabelSmoother.Params().Set(
      num_classes=vocab_size, uncertainty=ls_uncertainty)
  d.target_seq_len = 300
  d.beam_search.length_normalization = 0.2
  d.beam_search.coverage_penalty = 0.2
  p.decoder = d
  p.train.Set(learning_rate=learning_rate, l2_regularizer_weight=l2_regularizer_weight,
              clip_gradient_norm_to_value=5.0, grad_norm_to_clip_to_zero=100000.0,
              optimizer=optimizer.Adam.Params().Set(beta1=0.9, beta2=adam_beta2,
                                                    epsilon=adam_epsilon),
              lr_schedule=schedule.LinearRampupExponentialDecayScaledByNumSplitSchedule.Params(
              ).Set(warmup=lr_warmup_steps, decay_start=lr_decay_start,
                    decay_end=lr_decay_end, min=lr_min))
  p.eval.samples_per_summary = 12000
  return p


def SetupXEnDecTransformerParams(p, name, vocab_size, model_dim, hidden_dim, num_heads,
                                 num_layers, learning_rate, warmup_steps,
                                 residual_dropout_prob=0.1, input_dropout_prob=0.0,
                                 atten_dropout_prob=0.0, relu_dropout_prob=0.0,
                                 label_smoothing_uncertainty=0.1):
  """Fills a `TransformerXEnDecModel.Params()` (ref :593): X-encoder / X-decoder with the
  Transformer-base training recipe."""
  del atten_dropout_prob, relu_dropout_prob
  p.name = name
  enc = p.encoder
  enc.name = 'enc'
  enc.model_dim = model_dim
  enc.token_emb.Set(vocab_size=vocab_size, embedding_dim=model_dim,
                    params_init=py_utils.WeightInit.Gaussian(1.0 / model_dim ** 0.5))
  enc.position_emb.Set(embedding_dim=model_dim)
  enc.input_dropout_prob = input_dropout_prob
  enc.transformer_stack.Set(num_layers=num_layers, mdl_dim=model_dim, hidden_dim=hidden_dim,
                            num_atten_heads=num_heads, dropout_prob=residual_dropout_prob)
  dec = p.decoder
  dec.name = 'dec'
  dec.Set(source_dim=model_dim, model_dim=model_dim, num_trans_layers=num_layers,
          num_atten_heads=num_heads, hidden_dim=hidden_dim,
          input_dropout_prob=input_dropout_prob, per_word_avg_loss=True, target_seq_len=300)
  dec.token_emb.Set(vocab_size=vocab_size, embedding_dim=model_dim,
                    params_init=py_utils.WeightInit.Gaussian(1.0 / model_dim ** 0.5))
  dec.position_emb.Set(embedding_dim=model_dim)
  dec.softmax.Set(num_classes=vocab_size, input_dim=model_dim)
  if label_smoothing_uncertainty:
    dec.label_smoothing = layers.UniformLabelSmoother.Params().Set(
        num_classes=vocab_size, uncertainty=label_smoothing_uncertainty)
  p.train.Set(learning_rate=learning_rate, optimizer=optimizer.Adam.ParamsB(),
              clip_gradient_norm_to_value=0.0, grad_norm_to_clip_to_zero=0.0,
              lr_schedule=schedule.TransformerSchedule.Params().Set(
                  warmup_steps=warmup_steps, worker_replicas=1, model_dim=model_dim))
  p.eval.samples_per_summary = 12000
  return p
