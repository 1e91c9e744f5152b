"""Configuration mirror of the reference's yacs CfgNode for the inference path.

The reference reads ``model_config.yaml`` next to the checkpoint through yacs
(hamer/hamer/configs/__init__.py:88-113, hamer/hamer/models/__init__.py:32-47); that file is not in
the tree, the values below are those of configs_hydra/experiment/hamer_vit_transformer.yaml:25-43
and experiment/default.yaml:3-16.  yacs is not a dependency here: a small attribute dict is enough
for everything the path reads (cfg.MODEL.IMAGE_SIZE, IMAGE_MEAN/STD, BBOX_SHAPE, cfg.EXTRA.FOCAL_LENGTH,
cfg.MANO.*).
"""
import copy


class CfgNode(dict):
    """dict with attribute access and the yacs methods the reference calls (get/defrost/freeze)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def defrost(self):
        return self

    def freeze(self):
        return self

    def clone(self):
        return copy.deepcopy(self)


def _node(d):
    return CfgNode({k: _node(v) if isinstance(v, dict) else v for k, v in d.items()})


def default_config():
    return _node({
        'MODEL': {
            'IMAGE_SIZE': 256,
            'IMAGE_MEAN': [0.485, 0.456, 0.406],
            'IMAGE_STD': [0.229, 0.224, 0.225],
            'BBOX_SHAPE': [192, 256],              # forced by load_hamer (models/__init__.py:39-43)
            'BACKBONE': {'TYPE': 'vit'},
            'MANO_HEAD': {
                'TYPE': 'transformer_decoder', 'IN_CHANNELS': 2048,
                'TRANSFORMER_DECODER': {'depth': 6, 'heads': 8, 'mlp_dim': 1024, 'dim_head': 64,
                                        'dropout': 0.0, 'emb_dropout': 0.0, 'norm': 'layer',
                                        'context_dim': 1280},
            },
        },
        'MANO': {'DATA_DIR': '_DATA/data/', 'MODEL_PATH': '_DATA/data/mano', 'GENDER': 'neutral',
                 'NUM_HAND_JOINTS': 15, 'MEAN_PARAMS': '_DATA/data/mano_mean_params.npz',
                 'CREATE_BODY_POSE': False},
        'EXTRA': {'FOCAL_LENGTH': 5000, 'NUM_LOG_IMAGES': 4, 'NUM_LOG_SAMPLES_PER_IMAGE': 8,
                  'PELVIS_IND': 0},
    })
