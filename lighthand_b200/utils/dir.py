"""JSON dump with the reference's wire format (src/utils/dir.py:13-22): the payload is wrapped in a
one-element list; readers take ``meta[0]`` (src/utils/argparser.py:333,396)."""
import json


def dump(pred_out_path, xyz_pred_list):
    with open(pred_out_path, "w") as fo:
        json.dump([xyz_pred_list], fo)
