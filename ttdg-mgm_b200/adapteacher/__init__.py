"""B200-native drop-in for the reference's ``adapteacher`` package - test-time-adaptation hot path only
(SURVEY.md section 8).  Same module paths, class names, constructor signatures and state-dict keys as
/root/reference/adapteacher; every device op goes through libttdg_sm100.so (``ttdg_b200``)."""
