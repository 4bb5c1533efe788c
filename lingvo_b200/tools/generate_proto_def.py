r"""Writes `.proto` definitions for every wire format the in-repo protobuf codec speaks.

The reference ships `tools/generate_proto_def.cc`, which dumps TensorFlow's own descriptors
so that downstream builds can compile against them without a TF source tree. This
framework has no protobuf runtime at all (`utils/protowire.py` encodes / decodes by field
number), so the equivalent service is the other direction: emit schema files that describe
exactly the fields we read and write, for anyone who wants to consume our files with protoc
generated code (or to check wire compatibility with the reference's `*_pb2`).

  python -m lingvo_b200.tools.generate_proto_def /tmp/protos
"""

from __future__ import annotations

import os
import sys

_HEADER = 'syntax = "proto2";\n\n'

PROTOS = {
    'lingvo/core/ops/hyps.proto': _HEADER + '''package tensorflow.lingvo;

message Hypothesis {
  optional int32 beam_id = 1;
  repeated int32 ids = 2 [packed = true];
  repeated float scores = 3 [packed = true];
  message AttenVec {
    repeated float prob = 1 [packed = true];
  }
  repeated AttenVec atten_vecs = 4;
  optional float normalized_score = 5;
}
''',
    'lingvo/core/ops/versioned_file_set.proto': _HEADER + '''package tensorflow.lingvo;

message VersionedFileSet {
  optional FileSet current = 1;
  repeated FileSet history = 2;
}

message FileSet {
  repeated string file_pattern = 1;
  optional double create_timestamp = 2;
}
''',
    'tensorflow/core/example/example.proto': 'syntax = "proto3";\n\n' + '''package tensorflow;

message BytesList { repeated bytes value = 1; }
message FloatList { repeated float value = 1 [packed = true]; }
message Int64List { repeated int64 value = 1 [packed = true]; }
message Feature {
  oneof kind {
    BytesList bytes_list = 1;
    FloatList float_list = 2;
    Int64List int64_list = 3;
  }
}
message Features { map<string, Feature> feature = 1; }
message FeatureList { repeated Feature feature = 1; }
message FeatureLists { map<string, FeatureList> feature_list = 1; }
message Example { Features features = 1; }
message SequenceExample {
  Features context = 1;
  FeatureLists feature_lists = 2;
}
''',
    'tensorflow/core/util/event.proto': 'syntax = "proto3";\n\n' + '''package tensorflow;

message HistogramProto {
  double min = 1;
  double max = 2;
  double num = 3;
  double sum = 4;
  double sum_squares = 5;
  repeated double bucket_limit = 6 [packed = true];
  repeated double bucket = 7 [packed = true];
}
message Summary {
  message Image {
    int32 height = 1;
    int32 width = 2;
    int32 colorspace = 3;
    bytes encoded_image_string = 4;
  }
  message Value {
    string tag = 1;
    oneof value {
      float simple_value = 2;
      Image image = 4;
      HistogramProto histo = 5;
    }
  }
  repeated Value value = 1;
}
message Event {
  double wall_time = 1;
  int64 step = 2;
  oneof what {
    string file_version = 3;
    Summary summary = 5;
  }
}
''',
}


def Generate(output_dir: str):
  """Writes every schema below `output_dir`, creating directories; returns the paths."""
  written = []
  for rel, text in sorted(PROTOS.items()):
    path = os.path.join(output_dir, rel)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, 'w') as f:
      f.write(text)
    written.append(path)
  return written


def main(argv):
  if len(argv) != 2:
    print(__doc__)
    return 2
  for p in Generate(argv[1]):
    print(p)
  return 0


if __name__ == '__main__':
  sys.exit(main(sys.argv))
