"""Experiment configuration files: same syntax, keys and defaults as the reference
(config_utils.py:7-129), so scripts/config/*.config parse unchanged.

Syntax (config_utils.py:19-50): ``key = value`` lines, blank lines and lines starting with ``#``
ignored; a value containing ``[`` is a Python literal list; a value containing a digit and no
``/`` is a Python literal number; anything else is a raw string; a space inside a non-list
value is an error.  Deviation (documented reference defect): ``ctc_loss`` is
defaulted to 1 when absent (the reference tests the wrong key and would later KeyError).
"""
import ast
import os
import re
import sys

_LINE = re.compile(r'(\w+)\s*=\s*(.*)')


def load_configfile(cfile):
    """Parse a configuration file into a dict (config_utils.py:7-52)."""
    if not os.path.isfile(cfile):
        raise ValueError("Cannot find configuration file ", cfile)
    conf = {}
    with open(cfile, 'r') as fh:
        for nline, raw in enumerate(fh, start=1):
            line = raw.rstrip()
            if not line or line[0] == '#':
                continue
            par = _LINE.search(line)
            if par is None:
                raise ValueError("Wrong syntax in the configuration file at line ", nline)
            key, val = par.group(1), par.group(2)
            if '[' in val:
                try:
                    conf[key] = ast.literal_eval(val)
                except Exception:
                    raise ValueError("Wrong syntax in the configuration file at line {:d} "
                                     "(may be a missing square parenthesis?)".format(nline))
                continue
            if ' ' in val:
                raise ValueError("Wrong syntax in the configuration file at line {:d} "
                                 "(may be a space in the param value?)".format(nline))
            if re.search('[0-9]', val) and '/' not in val:
                try:
                    conf[key] = ast.literal_eval(val)
                except Exception:
                    raise ValueError("Wrong syntax in the configuration file at line {:d} "
                                     "(may be due to mixed letters and integers?)".format(nline))
            else:
                conf[key] = val
    return conf


def _default(config, key, value, warning):
    if key not in config:
        print("WARNING: " + warning, file=sys.stderr)
        config[key] = value


def check_trainconfiguration(config):
    """Validate / fill defaults (config_utils.py:55-129).  Call once per dict: like the reference
    it adds the CTC blank to ``num_asr_labels`` on every call."""
    if 'root_folder' not in config:
        raise ValueError("Root folder not defined")
    if 'exp_folder' not in config:
        raise ValueError("Experiment folder (exp_folder) not defined")
    config.setdefault('model_ckp', "")
    config.setdefault('model_ckp_vnet', "")
    _default(config, 'device', "/cpu:0", "using cpu as device has not been defined in the config file")
    if 'model' not in config:
        raise ValueError("Model type (model) not defined in config file")
    if 'net_dim' not in config:
        raise ValueError("Enhancement net dimensions (enh_net_dim) not defined in config file")
    _default(config, 'integration_layer', 0, "Embedding integration layer not defined in config file. Set to 0 by default")
    _default(config, 'audio_feat_dim', 257, "No. of audio input features of inpainting model not defined in config file. Set to 257 by default")
    _default(config, 'video_feat_dim', 136, "No. of video input features of inpainting model not defined in config file. Set to 136 by default")
    _default(config, 'audio_len', 16384, "Length of input wavs of inpainting model not defined in config file. Set to 0 by default (variable-length)")
    if 'audio_feat_mean' not in config:
        raise ValueError("File with mean of features (audio_feat_mean) not defined in config file")
    if 'audio_feat_std' not in config:
        raise ValueError("File with standard deviation of features (audio_feat_std) not defined in config file")
    _default(config, 'num_asr_labels', 33, "No. of speech recognition labels not defined in config file. Set to 33 by default")
    config['num_asr_labels'] += 1   # add the CTC "blank" label
    _default(config, 'ctc_loss', 1, "CTC loss weigth not defined in config file. Set to 1 by default")
    _default(config, 'batch_size', 1, "Batch size not defined in config file. Set to 1 by default")
    _default(config, 'dropout_rate', 0.0, "Dropout rate not defined in config file. Set to 1 by default")
    _default(config, 'starter_learning_rate', 0.06, "Starter learning rate not defined in config file. Set to 0.06 by default")
    _default(config, 'learning_rate', 0.06, "Learning rate not defined in config file. Set to 0.06 by default")
    _default(config, 'lr_updating_steps', 10000, "Updating steps of learning rate decay not defined in config file. Set to 10000 by default")
    _default(config, 'lr_decay', 0.5, "Learning rate decay not defined in config file. Set to 0.5 by default")
    _default(config, 'l2', 0.0, "L2 regularization coefficient not defined in config file. Set to 0 by default")
    _default(config, 'optimizer_type', 'adam', "Optimizer type not defined in config file. Set to 'adam' by default")
    if config['optimizer_type'] == 'momentum_dlr' and 'momentum' not in config:
        raise ValueError("momentum missing from config file")
    _default(config, 'max_n_epochs', 30, "max_n_epochs not defined. Set to 100 by default")
    _default(config, 'n_earlystop_epochs', 30, "n_earlystop_epochs not defined. Set to 3 by default")
    return config
