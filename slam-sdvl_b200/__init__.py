"""sdvl-b200: B200-native SDVL tracking front-end (package dir `slam-sdvl_b200/`, import name `slam_sdvl_b200`)."""
