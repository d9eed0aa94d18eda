"""get_package_share_directory('ros2_mpc') -> the read-only reference checkout, whose config/params.yaml the
reference's Mpc classes then load unchanged."""


def get_package_share_directory(name):
    assert name == "ros2_mpc"
    return "/root/reference"
