"""Image-text dual encoders for Crisscrossed Captions (ref
`lingvo/tasks/milan/params/cxc.py`)."""

from lingvo_b200 import model_registry
from lingvo_b200.models.milan import constants
from lingvo_b200.models.milan import image_preprocessor
from lingvo_b200.models.milan import tf_hub_layers
from lingvo_b200.models.milan import transformers
from lingvo_b200.models.milan.params import dual_encoder_recipe
from lingvo_b200.models.milan.params import generic_datasets

IMAGE = constants.Modality.IMAGE
TEXT = constants.Modality.TEXT


class _BaseImageTextRecipe(dual_encoder_recipe.DualEncoderRecipe):
  """ref :30."""

  def __init__(self):
    super().__init__()
    self.task_params.dual_encoder.loss_weights = {(IMAGE, TEXT): 0.5, (TEXT, IMAGE): 0.5}

  def AddEfficientNetB4ImageEncoder(self, image_feature='image/encoded', id_feature='image/id'):
    self.input_params.features_to_read += [image_feature, id_feature]
    self.AddPreprocessor(image_feature, image_preprocessor.ImagePreprocessor.Params().Set(
        output_image_size=tf_hub_layers.EFFICIENTNET_B4_INPUT_SHAPE))
    self.AddModality(IMAGE, input_features=image_feature, id_feature=id_feature,
                     encoder=tf_hub_layers.EfficientNetB4Params(),
                     output_dim=tf_hub_layers.EFFICIENTNET_B4_OUTPUT_FEATURE_DIM)

  def AddBertAdapterTextEncoder(self, bert_embeddings_feature='text/bert/token_features',
                                lengths_feature='text/bert/lengths', id_feature='text/id',
                                output_dim=768):
    self.input_params.features_to_read += [bert_embeddings_feature, lengths_feature, id_feature]
    input_dim = self.dataset.params.bert_dim
    self.AddModality(
        TEXT, input_features=(bert_embeddings_feature, lengths_feature), id_feature=id_feature,
        encoder=transformers.GetTransformerStackWithEmbeddingInput(
            input_dim=input_dim, num_layers=3, hidden_dim=3072, num_attention_heads=8,
            output_dim=output_dim),
        output_dim=output_dim)


@model_registry.RegisterSingleTaskModel
class EfficientNetB4BertAdapter(_BaseImageTextRecipe):
  """EfficientNet-B4-shaped image tower + adapter over pre-computed BERT features
  (ref :83)."""

  def __init__(self):
    super().__init__()
    self.AddEfficientNetB4ImageEncoder()
    self.AddBertAdapterTextEncoder()

  @property
  def default_dataset(self):
    return generic_datasets.ImageTextTFRecords.ParamsFromEnv().Instantiate()
