/* hw_stub.h -- a hosted stand-in for the parts of uc_tools' hw_stm32f103.h / libopencm3 that
 * stm32f103/mod_pdm_pwm.c and mod_controlrate.c touch, so that the REAL firmware sources compile
 * on the host and their timer ISR can be called in a loop (oracle/ref/ref_v2_isr.c).  Timer, GPIO
 * and NVIC calls do nothing; hw_multi_pwm_duty() -- the ISR's only output -- is captured; the
 * software interrupt that runs the control-rate update is delivered synchronously (on the MCU it
 * has the lower priority and runs when the PDM ISR returns; it touches line[1] and setpoint only,
 * the PDM ISR line[0] and the modulator state after the copy, so the order is immaterial).
 * Test infrastructure. */
#ifndef HW_STUB_H
#define HW_STUB_H
#include <stdint.h>
#define CONCAT_(a, b) a##b
#define CONCAT(a, b) CONCAT_(a, b)
#define ARRAY_SIZE(a) (sizeof(a) / sizeof((a)[0]))
#define infof(...) ((void)0)
enum { RCC_TIM3 = 1, RCC_GPIOA, RCC_GPIOB, TIM3, GPIOA, GPIOB, NVIC_TIM3_IRQ, HW_GPIO_CONFIG_ALTFN_2MHZ, HW_GPIO_CONFIG_OUTPUT };
struct hw_multi_pwm_gpio { uint32_t rcc, gpio, pin; };
struct hw_multi_pwm {                       /* field order of the initialiser at mod_pdm_pwm.c:59-64 */
    uint32_t rcc_tim, tim;
    struct hw_multi_pwm_gpio gpio[4];
    uint32_t gpio_config, div, duty, irq;
};
extern uint32_t ref_v2_duty[4];
#define hw_multi_pwm_init(c) ((void)(c))
#define hw_multi_pwm_start(c) ((void)(c))
#define hw_multi_pwm_stop(c) ((void)(c))
#define hw_multi_pwm_ack(c) ((void)(c))
#define hw_multi_pwm_duty(c, i, v) (ref_v2_duty[i] = (v))
static inline void hw_gpio_high(uint32_t port, uint32_t pin) { (void)port; (void)pin; }    /* PDM_CPU_USAGE_MARK is "GPIOA,3": two arguments */
static inline void hw_gpio_low(uint32_t port, uint32_t pin) { (void)port; (void)pin; }
static inline void hw_gpio_config(uint32_t port, uint32_t pin, uint32_t cfg) { (void)port; (void)pin; (void)cfg; }
#define HW_TIM_ISR(n) CONCAT(CONCAT(tim, n), _isr)
struct hw_swi { uint32_t line; };
#define HW_SWI_1 {1}
void exti1_isr(void);
#define hw_swi_init(c) ((void)(c))
#define hw_swi_ack(c) ((void)(c))
#define hw_swi_trigger(c) exti1_isr()
#endif
