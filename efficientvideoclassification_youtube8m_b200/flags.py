"""Flag side-channel of the reference models (tf.flags globals read inside create_model:
frame_level_models.py:15-47,218-219,237-238,286-289; video_level_models.py:13-19,421;
train.py:27-99, eval_finetune.py:21-60).  Same names; defaults are the reference's except where every run_*.sh
overrides them (`lstm_layers` 2 instead of 1, `feature_names`/`feature_sizes` "rgb, audio"/"1024, 128" instead of
"rgb"/"1024", `frame_features` True): the defaults here are the run_*.sh configuration.  Set them as attributes or
via ``parse``.  `sampling`, `precise` and `output_dir` are additions (BASELINE config #5; the converter's target directory)."""
from __future__ import annotations


class _Flags:
    def __init__(self):
        object.__setattr__(self, "_defs", {})

    def define(self, name, default, doc=""):
        self._defs[name] = (default, doc)
        object.__setattr__(self, name, default)

    def __setattr__(self, name, value):
        if name not in self._defs:
            raise AttributeError(f"unknown flag --{name}")
        object.__setattr__(self, name, value)

    def parse(self, argv):
        """--name value / --name=value pairs, as the run_*.sh launch lines pass them."""
        i = 0
        while i < len(argv):
            a = argv[i]
            if not a.startswith("--"):
                raise ValueError(f"unexpected argument {a}")
            if "=" in a:
                k, v = a[2:].split("=", 1)
                i += 1
            else:
                k, v = a[2:], argv[i + 1]
                i += 2
            if k not in self._defs:
                raise AttributeError(f"unknown flag --{k}")
            d = self._defs[k][0]
            if isinstance(d, bool):
                v = str(v).lower() in ("1", "true", "yes")
            elif isinstance(d, int):
                v = int(v)
            elif isinstance(d, float):
                v = float(v)
            object.__setattr__(self, k, v)

    def reset(self):
        for k, (d, _) in self._defs.items():
            object.__setattr__(self, k, d)


FLAGS = _Flags()
# frame_level_models.py:33-47
FLAGS.define("video_level_classifier_model", "MoeModel", "classifier applied to the final LSTM state")
FLAGS.define("lstm_cells", 1024, "Number of LSTM cells.")
FLAGS.define("lstm_layers", 2, "Number of LSTM layers (reference default 1; every run_*.sh passes 2).")
FLAGS.define("max_num_frames", 300, "maximum number of frames in a video")
FLAGS.define("num_inputs_to_lstm", 20, "number of chunks presented to the upper LSTM")
# video_level_models.py:14-16
FLAGS.define("moe_num_mixtures", 2, "The number of mixtures (excluding the dummy 'expert') used for MoeModel.")
# train.py:27-99
FLAGS.define("model", "HierarchicalLstmModel", "Which architecture to use for the model.")
FLAGS.define("label_loss", "CrossEntropyLoss", "Which loss function to use for training the model.")
FLAGS.define("optimizer", "AdamOptimizer", "What optimizer class to use.")
FLAGS.define("batch_size", 1024, "How many examples to process per batch for training.")
FLAGS.define("every_n", 1, "every nth frame to be used by student.")
FLAGS.define("dropout", 0.5, "Dropout Probability (unused by H-LSTM / MoE, SURVEY F15)")
FLAGS.define("regularization_penalty", 2.0, "weight of the regularization loss")
FLAGS.define("base_learning_rate", 0.001, "Which learning rate to start with.")
FLAGS.define("learning_rate_decay", 1.0, "Learning rate decay factor")
FLAGS.define("learning_rate_decay_examples", 4000000.0, "decay period in examples")
FLAGS.define("clip_gradient_norm", 1.0, "Norm to clip gradients to.")
FLAGS.define("top_k", 20, "How many predictions to output per video.")
FLAGS.define("feature_names", "rgb, audio", "features to use")
FLAGS.define("feature_sizes", "1024, 128", "lengths of the feature vectors")
# train.py:29-99 / eval_finetune.py:21-60 (launchers.py)
FLAGS.define("train_dir", "/tmp/yt8m_model/", "The directory to save the model files in / load them from.")
FLAGS.define("train_data_pattern", "", "File glob for the training dataset (tf.SequenceExample tfrecords).")
FLAGS.define("eval_data_pattern", "", "File glob defining the evaluation dataset.")
FLAGS.define("frame_features", True, "frame-level features (the H-LSTM path); run_*.sh pass True")
FLAGS.define("start_new_model", False, "If set, training does not resume from the latest checkpoint in train_dir.")
FLAGS.define("num_epochs", 10, "How many passes to make over the dataset before halting training.")
FLAGS.define("num_readers", 4, "How many threads to use for reading input files.")
FLAGS.define("gpu", 0, "GPU on which the code will run (single-process runs; torchrun sets LOCAL_RANK)")
FLAGS.define("run_once", False, "Whether to run eval only once.")
FLAGS.define("log_device_placement", False, "accepted for command-line parity, unused")
FLAGS.define("sampling", "uniform", "student frame sampler: uniform | random_frames | random_sequence")
FLAGS.define("precise", False, "split-bf16 operands (3 tensor-core products per contraction) instead of plain bf16")
FLAGS.define("output_dir", "", "train_convert_model: where the student-only checkpoint goes")
