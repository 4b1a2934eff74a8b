/* Stand-in for uc_tools fixedpoint.h (mod_pdm_pwm.c:11): nothing of it is used by the PDM ISR. */
