"""Headless runner: settings YAML reader, trajectory CSV writer, and an end-to-end run."""
import numpy as np
import pytest

from stereo_svo_slam_b200 import cli, synth

BLENDER_YAML = """%YAML:1.0
Camera1.fx: 470.0 # fx = 20mm/(32mm/752px) = 470.0 px
Camera1.fy: 470.0
Camera1.cx: 376.0
Camera1.cy: 240.0
Camera.baseline: 28.2 # 0.06 * 470px = 28.2
Camera1.k1: 0.0
Camera1.p2: 1.5e-3
Camera.grid_width: 75
Camera.grid_height: 48
Camera.search_x: 50
Camera.search_y: 6
Camera.window_size_pose_estimator: 4
Camera.window_size_opt_flow: 31
Camera.window_size_depth_calculator: 31
Camera.max_pyramid_levels: 5
Camera.min_pyramid_level_pose_estimation: 2
LEFT.K: !!opencv-matrix
   rows: 3
   data: [458.654, 0.0, 367.215]
"""


def test_read_settings_reference_keys(tmp_path):
    p = tmp_path / "Blender.yaml"
    p.write_text(BLENDER_YAML)
    s = cli.read_settings(str(p))
    assert s["fx"] == 470.0 and s["baseline"] == pytest.approx(28.2) and s["grid_width"] == 75 and s["grid_height"] == 48
    assert s["max_pyramid_levels"] == 5 and s["min_pyramid_level_pose_estimation"] == 2 and s["p2"] == pytest.approx(1.5e-3)
    assert s["k2"] == 0.0 and isinstance(s["search_x"], int)     # absent key reads as 0 (cv::FileStorage semantics)


def test_trajectory_csv_format(tmp_path):
    traj = np.array([[0, 0, 0, 0, 0, 0], [0.1, -0.2, 0.3, 0.02, -0.03, 0.01]], np.float32)
    p = tmp_path / "t.csv"
    cli.write_trajectory_csv(str(p), traj, [0.01, 0.025])
    rows = [list(map(float, l.split(","))) for l in p.read_text().strip().splitlines()]
    assert len(rows) == 2 and len(rows[0]) == 7 and rows[1][0] == pytest.approx(0.025)
    # slam_app.cpp:232-240: angles = Rodrigues((Ry*Rx)*Rz)
    R = (synth._rodrigues([0, -0.03, 0]) @ synth._rodrigues([0.02, 0, 0])) @ synth._rodrigues([0, 0, 0.01])
    assert np.allclose(synth._rodrigues(rows[1][4:7]), R, atol=1e-6)
    assert np.allclose(rows[1][1:4], [0.1, -0.2, 0.3], atol=1e-6)


def test_pairs_directory_reader(tmp_path):
    (tmp_path / "left").mkdir()
    (tmp_path / "right").mkdir()
    for k in range(3):
        np.save(tmp_path / "left" / f"{k:04d}.npy", np.full((4, 6), k, np.uint8))
        np.save(tmp_path / "right" / f"{k:04d}.npy", np.full((4, 6), 10 + k, np.uint8))
    got = list(cli.iter_pairs(str(tmp_path)))
    assert len(got) == 3 and got[2][0][0, 0] == 2 and got[2][1][0, 0] == 12 and got[1][2] == pytest.approx(0.05)


@pytest.mark.gpu
def test_cli_end_to_end_synthetic(tmp_path):
    out = tmp_path / "traj.csv"
    assert cli.main(["--synthetic", "S", "--frames", "12", "--trajectory", str(out)]) == 0
    rows = np.loadtxt(out, delimiter=",")
    assert rows.shape == (12, 7) and (np.diff(rows[:, 0]) > 0).all()
    seq = synth.make_sequence("S")
    assert np.abs(rows[-1, 1:4] - seq.pose(11)[:3]).max() < 0.03


def test_read_rectification_reference_yaml_layout(tmp_path):
    """LEFT.* / RIGHT.* !!opencv-matrix nodes in the layout of the reference's src/app/EuRoC.yaml (`data:[` and `data: [`)."""
    y = "%YAML:1.0\nCamera1.fx: 435.2\n"
    for side, k1 in (("LEFT", -0.2834), ("RIGHT", -0.2837)):
        y += f"{side}.height: 480\n{side}.width: 752\n"
        y += f"{side}.D: !!opencv-matrix\n   rows: 1\n   cols: 5\n   dt: d\n   data:[{k1}, 0.0739, 0.00019, 1.76e-05, 0.0]\n"
        y += f"{side}.K: !!opencv-matrix\n   rows: 3\n   cols: 3\n   dt: d\n   data: [458.654, 0.0, 367.215, 0.0, 457.296, 248.375, 0.0, 0.0, 1.0]\n"
        y += f"{side}.R:  !!opencv-matrix\n   rows: 3\n   cols: 3\n   dt: d\n   data: [1, 0, 0.008,\n 0, 1, 0.007, -0.008, -0.007, 1]\n"
        y += f"{side}.P:  !!opencv-matrix\n   rows: 3\n   cols: 4\n   dt: d\n   data: [435.2, 0, 367.45, -47.9,  0, 435.2, 252.2, 0,  0, 0, 1, 0]\n"
    p = tmp_path / "EuRoC.yaml"
    p.write_text(y)
    r = cli.read_rectification(str(p))
    assert r["LEFT"]["D"][0] == -0.2834 and r["RIGHT"]["D"][0] == -0.2837 and r["LEFT"]["K"].shape == (3, 3)
    assert r["RIGHT"]["P"].shape == (3, 4) and r["RIGHT"]["P"][0, 3] == -47.9 and r["LEFT"]["R"][2, 0] == -0.008
    assert r["LEFT"]["width"] == 752 and r["LEFT"]["height"] == 480
    (tmp_path / "bad.yaml").write_text(y.replace("RIGHT.K", "RIGHT.Kx"))
    with pytest.raises(ValueError, match="RIGHT.K"):
        cli.read_rectification(str(tmp_path / "bad.yaml"))


def test_iter_euroc_follows_data_csv(tmp_path):
    """EurocInput::load_images (euroc_input.cpp:87-116): rows of cam0/data.csv, '#' lines skipped, cam0 -> right, cam1 -> left,
    time = (ns - first ns) / 1e9 as float."""
    for cam, base in (("cam0", 10), ("cam1", 20)):
        (tmp_path / cam / "data").mkdir(parents=True)
        for k, name in enumerate(("1403636579763555584.npy", "1403636579813555456.npy")):
            np.save(tmp_path / cam / "data" / name, np.full((4, 6), base + k, np.uint8))
    (tmp_path / "cam0" / "data.csv").write_text("#timestamp [ns],filename\r\n1403636579763555584,1403636579763555584.npy\r\n"
                                                "1403636579813555456,1403636579813555456.npy\r\n\r\n")
    rows = list(cli.iter_euroc(str(tmp_path)))
    assert len(rows) == 2
    assert rows[0][0][0, 0] == 20 and rows[0][1][0, 0] == 10 and rows[0][2] == 0.0
    assert rows[1][0][0, 0] == 21 and rows[1][1][0, 0] == 11 and rows[1][2] == pytest.approx(0.049999872, abs=1e-6)
