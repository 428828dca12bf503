#ifndef GTCOMPAT_INIT_API_H
#define GTCOMPAT_INIT_API_H
#ifdef __cplusplus
extern "C" {
#endif
void gt_lib_init(void);
int gt_lib_clean(void);
#ifdef __cplusplus
}
#endif
#endif
