"""Entry point — `python3 vo_runner.py` from this directory, configured by config/vo_params.yaml exactly like
the reference's vo_runner.py (:6-16)."""
import numpy as np
import yaml

import vo_stereo_runner


def read_yaml_file():
    with open("config/vo_params.yaml") as f:
        params = yaml.load(f, Loader=yaml.FullLoader)
    cam_intr = np.reshape(np.asarray(params["camera_intrinsic_matrix"], dtype=np.float64), (3, 3))
    return vo_stereo_runner.vo_offline_data(cam_intr, params["image_path"], params["output_filename"])


if __name__ == "__main__":
    read_yaml_file()
