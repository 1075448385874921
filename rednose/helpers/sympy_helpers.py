from rednose_b200.geometry import (cross, euler2quat, euler2rot, euler_rotate, quat2rot, quat_matrix_l,  # noqa: F401
                                   quat_matrix_r, quat_rotate, rot_matrix, rot_to_euler, rotations_from_quats)
