// autoptr.h -- intrusive-count shared pointer with the reference's public layout (include/gmapping/utils/autoptr.h:12-77):
// consumers reach into m_reference->data / ->shares directly (e.g. the MAP_CONSISTENCY_CHECK code), so the member
// names are part of the API.  Converts to int: non-zero when it points at something.
#ifndef GMB200_AUTOPTR_H
#define GMB200_AUTOPTR_H
#include <assert.h>

namespace GMapping {

template <class X>
class autoptr {
  public:
    struct reference {
        X* data;
        unsigned int shares;
    };
    inline autoptr(X* p = (X*)(0)) : m_reference(0) {
        if (p) {
            m_reference = new reference;
            m_reference->data = p;
            m_reference->shares = 1;
        }
    }
    inline autoptr(const autoptr<X>& ap) : m_reference(ap.m_reference) {
        if (m_reference) m_reference->shares++;
    }
    inline autoptr& operator=(const autoptr<X>& ap) {
        if (m_reference == ap.m_reference) return *this;
        release();
        m_reference = ap.m_reference;
        if (m_reference) m_reference->shares++;
        return *this;
    }
    inline ~autoptr() { release(); }
    inline operator int() const { return m_reference && m_reference->shares && m_reference->data; }
    inline X& operator*() {
        assert(m_reference && m_reference->shares && m_reference->data);
        return *(m_reference->data);
    }
    inline const X& operator*() const {
        assert(m_reference && m_reference->shares && m_reference->data);
        return *(m_reference->data);
    }
    reference* m_reference;

  private:
    inline void release() {
        if (m_reference && --m_reference->shares == 0) {
            delete m_reference->data;
            delete m_reference;
        }
        m_reference = 0;
    }
};

}  // namespace GMapping
#endif
