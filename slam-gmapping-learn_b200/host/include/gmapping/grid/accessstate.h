// accessstate.h (reference API: include/gmapping/grid/accessstate.h)
#ifndef GMB200_ACCESSSTATE_H
#define GMB200_ACCESSSTATE_H
namespace GMapping {
enum AccessibilityState { Outside = 0x0, Inside = 0x1, Allocated = 0x2 };
}
#endif
